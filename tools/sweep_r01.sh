python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in 0 1 2 3; do
  FD_K6_CFG=$cfg python bench.py --steps 3 --warmup 3 > gpurun_out/sw_k6_$cfg.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/sw_k6_$cfg.json'));print('K6_CFG=$cfg value',round(d['value']),'e2e',round(d['e2e']['value']),'verify_ms',round(d['stages_ms_per_step']['verify'],2),'scan_ms',round(d['stages_ms_per_step']['scan'],3), d['search_wall_ms'])"
done
for kb in 36 110 220; do
  FD_K3_TILE_KB=$kb python bench.py --steps 3 --warmup 3 > gpurun_out/sw_k3_$kb.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/sw_k3_$kb.json'));print('K3_TILE_KB=$kb scan_ms',round(d['stages_ms_per_step']['scan'],3),'select',round(d['stages_ms_per_step']['select'],3))"
done
