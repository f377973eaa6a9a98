#!/bin/bash
# Round-2 call 18 (1 GPU): where the ring forward's time goes at d = 128 (parts disabled one at a time), then one ncu capture.
set -u
O=gpurun_out; mkdir -p $O
for dbg in 0 1 2 4 8 16 3 5 7 15; do echo "dbg=$dbg"; YR_FWD_DBG=$dbg timeout 120 python scripts/dense_bench.py fwd 2>&1 | grep "mode=1" | grep -v "n=4099"; done > $O/r02_dense_fwd_dbg.txt 2>&1
cat $O/r02_dense_fwd_dbg.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ngcf_dense_fwd_tc_kernel -s 14 -c 2 -o $O/r02_fwd_ring_d128 -f python scripts/dense_bench.py fwd > $O/ncu_fwd.log 2>&1; echo "ncu rc=$?"
