#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/probe_e2e_threads.py > gpurun_out/r7_probe.jsonl 2> gpurun_out/r7_probe.err
tail -3 gpurun_out/r7_probe.err
cat gpurun_out/r7_probe.jsonl
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"
