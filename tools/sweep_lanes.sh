python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for l in 1 2 3 4; do
  FD_VERIFY_LANES=$l python bench.py --steps 5 --warmup 3 --sweep "" > gpurun_out/sw_l$l.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/sw_l$l.json'));print('LANES=$l value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],2),round(d['e2e']['ms_per_step'],2),d['search_wall_ms'])"
done
