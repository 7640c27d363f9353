#!/bin/bash
# round 2, call 2: line-flush version of the ring partition path: parity, A/B, ncu
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or tile or headline or bin_edges" > gpurun_out/r3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3_pytest.log
tail -5 gpurun_out/r3_pytest.log
for cfg in "ring 1" "ring 2" "tile 1"; do
  set -- $cfg
  B200_COUNT_PATH=$1 B200_RING_FG=$2 python tools/ab_headline.py --rows 1e9 --reps 10 --tag "$1-fg$2" >> gpurun_out/r3_ab.jsonl 2>gpurun_out/r3_ab_err.log
done
cat gpurun_out/r3_ab.jsonl
B200_RING_FG=1 ncu --set full --clock-control none --import-source on -k regex:k_ring -s 4 -c 2 -o gpurun_out/r3_ring python tools/ab_headline.py --rows 1e9 --reps 1 > gpurun_out/r3_ncu.log 2>&1
ls -la gpurun_out
