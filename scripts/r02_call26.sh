#!/bin/bash
# Round-2 call 26 (1 GPU): does the round-2 base (commit b2d83f8) meet 1e-5 on the sharded d = 128 Adam case, and how much does it vary run to run?
set -u
O=gpurun_out; mkdir -p $O
cd old_tree
for i in 1 2 3; do timeout -s KILL 200 python scripts/adam_dense_modes.py 2>&1 | grep "mode=1" ; done > ../$O/r02_adam_modes_old.txt 2>&1
cd ..
cat $O/r02_adam_modes_old.txt
for i in 1 2; do timeout -s KILL 200 python scripts/adam_dense_modes.py 2>&1 | grep "d=128 mode=0" ; done > $O/r02_adam_modes_new0.txt 2>&1
cat $O/r02_adam_modes_new0.txt
