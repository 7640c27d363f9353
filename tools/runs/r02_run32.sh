#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "ring or parity or golden or frame and not sweep_all" 2>&1 | tail -3
S=$PWD/vaex_b200/libb200agg_static.so
(for i in 1 2; do
python tools/ab_headline.py --reps 20 --tag tickets
VAEX_B200_LIB=$S python tools/ab_headline.py --reps 20 --tag static
done
python tools/ab_headline.py --reps 20 --tag tickets --occupy 16,120
VAEX_B200_LIB=$S python tools/ab_headline.py --reps 20 --tag static --occupy 16,120
python tools/ab_headline.py --reps 20 --tag tickets --occupy 32,200
VAEX_B200_LIB=$S python tools/ab_headline.py --reps 20 --tag static --occupy 32,200
python tools/ab_headline.py --reps 20 --tag tickets --rows 2.5e8
VAEX_B200_LIB=$S python tools/ab_headline.py --reps 20 --tag static --rows 2.5e8
) > gpurun_out/r32_ab.jsonl 2> gpurun_out/r32_ab.err
tail -3 gpurun_out/r32_ab.err; cut -c1-330 gpurun_out/r32_ab.jsonl
