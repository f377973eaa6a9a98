#!/bin/bash
# Round-2 call 42 (1 GPU): the headline part of the bench line on the final tree (no extras, no config 5, no CPU legs).
set -u
O=gpurun_out; mkdir -p $O
timeout -s KILL 120 python bench.py --no-extra --no-c5 --no-cpu-baseline --steps 20 --warmup 5 > $O/r02_bench_headline.json 2> $O/r02_bench_headline.err; echo "bench rc=$?"; tail -c 700 $O/r02_bench_headline.json
