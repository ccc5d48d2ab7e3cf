#!/bin/bash
# quick GPU pass: verification / e2e parity tests and a short bench line (gpurun -- bash tools/gpu_quick.sh TAG [pytest -k expr])
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_parity.py -m gpu -x -q ${2:+-k "$2"} > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --sweep "" --cpu-sample 64 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"e2e_ms",round(d["e2e"]["ms_per_step"],3))
print("stages",{k:round(v,3) for k,v in d["stages_ms_per_step"].items()})
print("wall",d["search_wall_ms"],"prep",d["e2e_prepare_host_ms"])
print("parity",d["parity_check"]["mismatches"],d["parity_check"]["rows_checked"],"roofline",d["roofline"]["frac"],d["roofline"]["launch_ms"])
print("shipped",d.get("shipped_motifs"))
PY
