#!/bin/bash
# second ncu pass of round 2: launch list of the bench default restricted to this library's kernels, and the Transform
# kernel after the change to 4 documents per warp
set -u
O=gpurun_out/r2
mkdir -p $O
A="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-parity --skip-extras"
C="python bench.py --workload C5 --docs 524288 --steps 1 --warmup 1"
KERN='regex:scvb0|phi_blend|hat_colsum|nz_update|sum_partials|rho_table|colsum|perplexity|memory_mix|transpose|narrow|doc_wc|pcg_fill|inv_z|reciprocal'
$A > $O/plain_c3b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 400 --csv --log-file $O/launches_c3b.csv $A > $O/ncu_launches_c3b.log 2>&1
echo "launch list rc=$?"
$C > $O/plain_c5b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scvb0_docsg -s 8 -c 1 -o $O/prof_c5b $C > $O/ncu_full_c5b.log 2>&1
echo "c5 rc=$?"
