#!/bin/bash
mkdir -p gpurun_out
(echo "== NT copy, 4MB x4"; python tools/probe_e2e_threads.py 28 8,16,32 20,24
echo "== NT copy, 2MB x8"; B200_BOUNCE_PIECE_KB=2048 B200_BOUNCE_COUNT=8 python tools/probe_e2e_threads.py 28 16 20,24
echo "== memcpy, 4MB x4"; B200_BOUNCE_MEMCPY=1 python tools/probe_e2e_threads.py 28 16 20,24
) > gpurun_out/r19_probe.txt 2>&1
cat gpurun_out/r19_probe.txt
python -m pytest tests -x -q -m gpu -k "not sweep_all" 2>&1 | tail -3
