#!/bin/bash
# full GPU test suite, bench line, and a full ncu capture of the verification kernels
TAG=${1:-v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --sweep "" --cpu-sample 64 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k6b_components|k6a_table|k6c_kabsch' -c 9 \
  -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 1 --sweep "" --shipped 0 --parity-queries 1 --cpu-sample 16 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"e2e_ms",round(d["e2e"]["ms_per_step"],3))
print("stages",{k:round(v,3) for k,v in d["stages_ms_per_step"].items()})
print("wall",d["search_wall_ms"],"prep",d["e2e_prepare_host_ms"])
print("parity",d["parity_check"]["mismatches"],d["parity_check"]["rows_checked"],"roofline",d["roofline"]["frac"],d["roofline"]["launch_ms"])
print("index",{k:v for k,v in d["index_build"].items() if k!="note"})
PY
