#!/bin/bash
# K3 alone: parity tests, variants of the scan knobs at 1x and 4x the index, then a full ncu capture of k3_scan_v3
TAG=${1:-k3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
V=${2:-"FD_K3_RUNS=2,FD_K3_DWARPS=16,FD_K3_CTAS=1,FD_K3_CTAS=1+FD_K3_THREADS=1024+FD_K3_DWARPS=16,FD_K3_LIMIT=0"}
timeout 600 python tools/k3_probe.py --distinct --steps 5 --variants "$V" > gpurun_out/${TAG}_probe1.jsonl 2> gpurun_out/${TAG}_probe1.err; echo "probe1 rc=$?"
timeout 900 python tools/k3_probe.py --distinct --steps 3 --tile 4 --variants "$V" > gpurun_out/${TAG}_probe4.jsonl 2> gpurun_out/${TAG}_probe4.err; echo "probe4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_scan_v3' -c 2 -o gpurun_out/${TAG}_full -f python tools/k3_probe.py --distinct --steps 1 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_probe1.jsonl","gpurun_out/${TAG}_probe4.jsonl"):
    for l in open(f):
        d=json.loads(l); print(d["variant"], d["structures"], "scan", round(d["scan"],3), "select", round(d["select"],3), "GB/s", round(d["scan_GBps"],1), "rows", d["rows"])
PY
