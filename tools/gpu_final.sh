#!/bin/bash
# final GPU pass of a round: the whole GPU test suite, smoke, one bench line (gpurun -- bash tools/gpu_final.sh TAG)
TAG=${1:-final}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest.log | head -20
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"parity",d["parity_check"]["mismatches"],"roofline",d["roofline"]["frac"])
print("widened",json.dumps(d.get("widened_rows"))[:3000])
PY
