#!/bin/bash
# One GPU-box pass: parity tests, both bench arms, ncu launch list and full captures (gpurun -- bash tools/gpu_round.sh TAG)
TAG=${1:-r03a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --sweep "4,23" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --sweep "" --shipped 0 --parity-queries 4 --cpu-sample 16 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k3_scan_v3|k6a_table|k6b_components|k6c_kabsch|k6d_rows|k3_topn_sort|k3_lookup' -c 16 \
  -o gpurun_out/${TAG}_full -f python bench.py --steps 1 --warmup 1 --sweep "" --shipped 0 --parity-queries 1 --cpu-sample 16 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
head -c 3000 gpurun_out/${TAG}_bench.json
