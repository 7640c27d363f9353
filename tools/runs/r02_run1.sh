#!/bin/bash
# round 2, call 1: parity of the new ring partition path + A/B against round 1's tile path + ncu of the new pair
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or tile or headline or bin_edges" > gpurun_out/r1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
tail -5 gpurun_out/r1_pytest.log
for cfg in "ring 1" "ring 2" "tile 1"; do
  set -- $cfg
  B200_COUNT_PATH=$1 B200_RING_FG=$2 python tools/ab_headline.py --rows 1e9 --reps 10 --tag "$1-fg$2" >> gpurun_out/r1_ab.jsonl 2>gpurun_out/r1_ab_err.log
done
cat gpurun_out/r1_ab.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ring -c 8 --csv --log-file gpurun_out/r1_launches.csv python tools/ab_headline.py --rows 1e9 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ring -s 4 -c 2 -o gpurun_out/r1_ring python tools/ab_headline.py --rows 1e9 --reps 1 > gpurun_out/r1_ncu.log 2>&1
ls -la gpurun_out
