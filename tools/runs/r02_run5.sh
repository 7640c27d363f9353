#!/bin/bash
# round 2, call 5: keys through registers (no TMA staging) vs TMA staging, 128-byte vs 64-byte lines
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or tile or headline or bin_edges" > gpurun_out/r5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r5_pytest.log
VAEX_B200_LIB=$PWD/vaex_b200/libb200agg_reg32.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or tile" >> gpurun_out/r5_pytest.log 2>&1
echo "pytest(reg32) rc=$?" >> gpurun_out/r5_pytest.log
tail -8 gpurun_out/r5_pytest.log
python tools/ab_headline.py --rows 1e9 --reps 10 --tag "reg-line64" >> gpurun_out/r5_ab.jsonl 2>gpurun_out/r5_ab_err.log
VAEX_B200_LIB=$PWD/vaex_b200/libb200agg_reg32.so python tools/ab_headline.py --rows 1e9 --reps 10 --tag "reg-line32" >> gpurun_out/r5_ab.jsonl 2>>gpurun_out/r5_ab_err.log
VAEX_B200_LIB=$PWD/vaex_b200/libb200agg_tma64.so python tools/ab_headline.py --rows 1e9 --reps 10 --tag "tma-line64" >> gpurun_out/r5_ab.jsonl 2>>gpurun_out/r5_ab_err.log
B200_RING_NO_CLAMP=1 python tools/ab_headline.py --rows 1e9 --reps 10 --tag "reg-line64-noclamp" >> gpurun_out/r5_ab.jsonl 2>>gpurun_out/r5_ab_err.log
cat gpurun_out/r5_ab.jsonl
ncu --set full --clock-control none --import-source on -k regex:k_ring -s 4 -c 2 -o gpurun_out/r5_ring python tools/ab_headline.py --rows 1e9 --reps 1 > gpurun_out/r5_ncu.log 2>&1
ls -la gpurun_out | tail -8
