#!/bin/bash
# N-GPU bench line + the 2-rank NCCL tests (gpurun --gpus N -- bash tools/gpu_ngpu.sh TAG N)
TAG=${1:-n2}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n$N.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"e2e_ms",round(d["e2e"]["ms_per_step"],3))
print("stages",{k:round(v,3) for k,v in d["stages_ms_per_step"].items()})
print("wall",d["search_wall_ms"],"prep",d["e2e_prepare_host_ms"])
print("parity",d["parity_check"]["mismatches"],d["parity_check"]["rows_checked"],"roofline",d["roofline"]["frac"],d["roofline"]["launch_ms"],d["roofline"]["launch_ms_max_over_ranks"])
print("exchange",d.get("exchange"))
PY
