#!/bin/bash
# full GPU test suite + bench line (optionally with the Swiss-Prot-scale scan sweep: SWEEP="4,23")
TAG=${1:-f}; SWEEP=${2:-4}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 1500 python bench.py --steps 5 --warmup 3 --sweep "$SWEEP" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"e2e_ms",round(d["e2e"]["ms_per_step"],3))
print("stages",{k:round(v,3) for k,v in d["stages_ms_per_step"].items()})
print("parity",d["parity_check"]["mismatches"],d["parity_check"]["rows_checked"],"roofline",d["roofline"]["frac"],d["roofline"]["launch_ms"])
for r in d.get("scan_vs_index_size",[]): print(r)
print("cpu", d.get("cpu_baseline"))
PY
