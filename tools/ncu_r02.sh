#!/bin/bash
# ncu captures of round 2 (run under gpurun on one B200): launch list + full set of the minibatch kernel for the bench
# default (configs[2]), the K = 1024 fit (configs[3]) and the Transform kernel (configs[4], 524,288 documents)
set -u
O=gpurun_out/r2
mkdir -p $O
A="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-parity --skip-extras"
B="python bench.py --workload C4 --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-parity --skip-extras"
C="python bench.py --workload C5 --docs 524288 --steps 1 --warmup 1"
$A > $O/plain_c3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv $A > $O/ncu_launches_c3.log 2>&1
echo "launch list rc=$?"
$A > $O/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"scvb0_docs32|phi_blend|hat_colsum" -s 6 -c 4 -o $O/prof_c3 $A > $O/ncu_full_c3.log 2>&1
echo "c3 rc=$?"
$B > $O/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scvb0_docs32 -s 4 -c 2 -o $O/prof_c4 $B > $O/ncu_full_c4.log 2>&1
echo "c4 rc=$?"
$C > $O/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scvb0_docs32 -s 8 -c 2 -o $O/prof_c5 $C > $O/ncu_full_c5.log 2>&1
echo "c5 rc=$?"
ls -la $O/*.ncu-rep
