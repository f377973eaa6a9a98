#!/bin/bash
# Round-1 ncu evidence (run under gpurun). Every ncu run is preceded by the same command without ncu.
set -u
O=gpurun_out
NG="python bench.py --only ngcf --steps 2 --warmup 1"
EV="python bench.py --only eval --steps 1 --warmup 1"
MF="python bench.py --only mf --steps 50 --warmup 3"
$NG > $O/plain_ngcf.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_ngcf.csv $NG > $O/ncu_ngcf.log 2>&1
$NG > $O/plain_ngcf2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"spmm_chunk|dense_fwd|dense_bwd|ngcf_tail|dense_opt" -s 16 -c 16 -o $O/prof_ngcf $NG > $O/ncu_ngcf_full.log 2>&1
$EV > $O/plain_eval.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:eval_topk -s 1 -c 1 -o $O/prof_eval $EV > $O/ncu_eval_full.log 2>&1
$MF > $O/plain_mf.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bpr_mf_train -s 1 -c 1 -o $O/prof_mf $MF > $O/ncu_mf_full.log 2>&1
tail -2 $O/plain_ngcf.log $O/plain_eval.log $O/plain_mf.log
tail -3 $O/ncu_ngcf_full.log $O/ncu_eval_full.log $O/ncu_mf_full.log
ls -la $O
