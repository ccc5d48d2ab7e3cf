#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run on the GPU box: gpurun -- bash tools/sanitize.sh).
# r01 result (one B200): memcheck 0 errors on test_gpu_e2e (config 1, 600-structure synthetic pipeline, search_stream),
# test_gpu_parity and test_gpu_sharded; racecheck 0 hazards on the config-1 and synthetic e2e tests.
# Round 2 selection: the kernels written this round -- k3_scan_v3 (runs of tiles, warp queues, red.shared votes),
# k3w_* (whole-structure path), k6b (components shared by the warps of a CTA), k6d_rows, the amino-acid directory.
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_e2e.py -x -q \
  -k "config1 or 600-5-None-0 or repeated_batch or whole_structure" > gpurun_out/sanitize_memcheck_e2e.log 2>&1; echo "memcheck e2e rc=$?"
tail -3 gpurun_out/sanitize_memcheck_e2e.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q \
  -k "700-21 or whole_structure or id_range or candidate_edges" > gpurun_out/sanitize_memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"
tail -3 gpurun_out/sanitize_memcheck_parity.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_e2e.py -x -q \
  -k "config1 or 600-5-None-0" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/sanitize_racecheck.log
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_*.log | sort | uniq -c
