#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 99 --launch-timeout 0 \
  python -m pytest tests -x -q -m gpu -k "ring_partition_variants or agg_list or test_gpu_hash or strings or smoke" > gpurun_out/r28_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r28_racecheck.log
grep -c "hazard" gpurun_out/r28_racecheck.log
grep -m 12 -A6 "hazard" gpurun_out/r28_racecheck.log | cut -c1-220
tail -8 gpurun_out/r28_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 99 python -m pytest tests -x -q -m gpu -k "ring_partition_variants or agg_list or test_gpu_hash" > gpurun_out/r28_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/r28_synccheck.log; tail -4 gpurun_out/r28_synccheck.log
