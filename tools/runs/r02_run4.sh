#!/bin/bash
# round 2, call 4: 64-byte lines (24 warps/SM) vs 128-byte lines, clamped magic-floor index vs F2I index
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or tile or headline or bin_edges" > gpurun_out/r4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r4_pytest.log
B200_RING_NO_CLAMP=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or tile" >> gpurun_out/r4_pytest.log 2>&1
echo "pytest(noclamp) rc=$?" >> gpurun_out/r4_pytest.log
VAEX_B200_LIB=$PWD/vaex_b200/libb200agg_line64.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring or tile" >> gpurun_out/r4_pytest.log 2>&1
echo "pytest(line64) rc=$?" >> gpurun_out/r4_pytest.log
tail -12 gpurun_out/r4_pytest.log
python tools/ab_headline.py --rows 1e9 --reps 10 --tag "line32-clamp" >> gpurun_out/r4_ab.jsonl 2>gpurun_out/r4_ab_err.log
B200_RING_NO_CLAMP=1 python tools/ab_headline.py --rows 1e9 --reps 10 --tag "line32-noclamp" >> gpurun_out/r4_ab.jsonl 2>>gpurun_out/r4_ab_err.log
VAEX_B200_LIB=$PWD/vaex_b200/libb200agg_line64.so python tools/ab_headline.py --rows 1e9 --reps 10 --tag "line64-clamp" >> gpurun_out/r4_ab.jsonl 2>>gpurun_out/r4_ab_err.log
VAEX_B200_LIB=$PWD/vaex_b200/libb200agg_line64.so B200_RING_NO_CLAMP=1 python tools/ab_headline.py --rows 1e9 --reps 10 --tag "line64-noclamp" >> gpurun_out/r4_ab.jsonl 2>>gpurun_out/r4_ab_err.log
cat gpurun_out/r4_ab.jsonl
ncu --set full --clock-control none --import-source on -k regex:k_ring -s 4 -c 2 -o gpurun_out/r4_ring python tools/ab_headline.py --rows 1e9 --reps 1 > gpurun_out/r4_ncu.log 2>&1
ls -la gpurun_out | tail -8
