#!/bin/bash
# First GPU pass over what the last session of round 2 wrote without a GPU (gpurun -- bash tools/gpu_r06.sh TAG):
# the two new GPU tests, then the bench line with and without the serving loop.
TAG=${1:-r06a}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zz_serving.py tests/test_gpu_zzz_foldcomp.py -m gpu -q --timeout 240 -p no:cacheprovider \
    > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -5 gpurun_out/${TAG}_pytest_new.log
timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --shipped 0 --sweep '' > gpurun_out/${TAG}_bench.json \
    2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --shipped 0 --sweep '' --pipeline 0 > gpurun_out/${TAG}_bench_nopipe.json \
    2> gpurun_out/${TAG}_bench_nopipe.err; echo "bench (no serving loop) rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_nopipe.json"):
    d = json.load(open(f))
    print(f, "value", round(d["value"]), "e2e", json.dumps(d["e2e"])[:700], "prep", d["e2e_prepare_host_ms"], "parity",
          d["parity_check"]["mismatches"])
PY
