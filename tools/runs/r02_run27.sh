#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
# memcheck over the kernels added / rewritten this round (small cases; the sweep and the > 2^30-row test are left out)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 0 --target-processes all \
  python -m pytest tests -x -q -m gpu -k "ring_partition_variants or agg_list or strings or minmax or expr or filter or combine or protocol or hash and not sweep_all and not more_than_one_batch" \
  > gpurun_out/r27_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r27_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r27_memcheck.log
tail -25 gpurun_out/r27_memcheck.log
