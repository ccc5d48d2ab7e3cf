#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run on the GPU box: gpurun -- bash tools/sanitize.sh).
# r01 result (one B200): memcheck 0 errors on test_gpu_e2e (config 1, 600-structure synthetic pipeline, search_stream),
# test_gpu_parity and test_gpu_sharded; racecheck 0 hazards on the config-1 and synthetic e2e tests
# (shared-memory queues of k6a, vote tiles of k3_scan, warp state of k6b).
set -e
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_e2e.py -x -q -k "config1 or 600-5-None-0 or stream"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_e2e.py -x -q -k "config1 or 600-5-None-0"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -x -q
