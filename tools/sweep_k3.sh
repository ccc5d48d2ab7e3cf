for kb in 72 150 220; do for th in 256 512 1024; do
  FD_K3_TILE_KB=$kb FD_K3_THREADS=$th python bench.py --steps 3 --warmup 2 --sweep "" > gpurun_out/sw.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/sw.json'));print('TILE_KB=$kb THREADS=$th scan_ms',round(d['stages_ms_per_step']['scan'],3),'select',round(d['stages_ms_per_step']['select'],3),'value',round(d['value']))"
done; done
