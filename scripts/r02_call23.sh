#!/bin/bash
# Round-2 call 23 (1 GPU): bulk L2 prefetch distance (0 / 1 / 2 / 3 tiles) for the ring forward / backward at d = 128: time and DRAM bytes.
set -u
O=gpurun_out; mkdir -p $O
for pf in 0 1 2 3; do echo "fwd pf=$pf"; YR_FWD_DBG=$((pf*256)) timeout -s KILL 120 python scripts/dense_bench.py fwd 2>&1 | grep "mode=1" | grep -v "n=4099"; done > $O/r02_dense_pf.txt 2>&1
for pf in 0 1 2 17 18; do echo "bwd pf=$pf"; YR_BWD_DBG=$pf timeout -s KILL 120 python scripts/dense_bench.py bwd 2>&1 | grep "mode=2" | grep -v "n=4099"; done >> $O/r02_dense_pf.txt 2>&1
cat $O/r02_dense_pf.txt
for pf in 0 1; do
YR_FWD_DBG=$((pf*256)) timeout -s KILL 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ngcf_dense_fwd_tc_kernel -s 72 -c 1 python scripts/dense_bench.py fwd 2>&1 | grep -E "dram__|gpu__time" | sed "s/^/fwd pf=$pf /"
YR_BWD_DBG=$pf timeout -s KILL 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ngcf_dense_bwd_tc_kernel -s 74 -c 1 python scripts/dense_bench.py bwd 2>&1 | grep -E "dram__|gpu__time" | sed "s/^/bwd pf=$pf /"
done > $O/r02_dense_pf_dram.txt 2>&1
cat $O/r02_dense_pf_dram.txt
