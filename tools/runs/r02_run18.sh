#!/bin/bash
mkdir -p gpurun_out
(echo "== default 4MB x4"; python tools/probe_e2e_threads.py 28 4,8,16,32 20,22,24
echo "== 1MB x8"; B200_BOUNCE_PIECE_KB=1024 B200_BOUNCE_COUNT=8 python tools/probe_e2e_threads.py 28 8,16 20,24
echo "== 16MB x4"; B200_BOUNCE_PIECE_KB=16384 B200_BOUNCE_COUNT=4 python tools/probe_e2e_threads.py 28 8,16 20,24
echo "== 64MB x2"; B200_BOUNCE_PIECE_KB=65536 B200_BOUNCE_COUNT=2 python tools/probe_e2e_threads.py 28 8,16 20,24
lscpu | grep -i "numa\|model name\|socket\|L3"; nvidia-smi topo -m 2>/dev/null | head -20) > gpurun_out/r18_probe.txt 2>&1
cat gpurun_out/r18_probe.txt
