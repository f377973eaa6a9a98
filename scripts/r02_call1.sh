#!/bin/bash
# Round-2 call 1: measured L2 gather ceiling + SpMM cache/occupancy counters.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r02_gpu.txt 2>&1
timeout 120 scripts/bin/l2_gather_bench > $O/r02_l2_gather.txt 2>&1; echo "l2bench rc=$?"
timeout 300 python scripts/spmm_bench.py 50 > $O/r02_spmm_plain.txt 2>&1; echo "spmm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum.per_second,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg,lts__t_bytes.sum.per_second \
  --clock-control none -k regex:spmm_chunk -c 12 --csv --log-file $O/r02_spmm_counters.csv python scripts/spmm_bench.py 2 > $O/r02_spmm_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02_tests1.log 2>&1; echo "tests rc=$?"; tail -4 $O/r02_tests1.log
cat $O/r02_l2_gather.txt; tail -5 $O/r02_spmm_plain.txt
