#!/bin/bash
# third ncu pass of round 2: the paired (two tokens per butterfly) fit kernel against the unpaired one on a launch range of
# 1,600 documents (BatchSize 100 x Processes 16, the reference's defaults on a 16-thread host), and the launch list of one
# iteration at that operating point.  Each command first runs plainly.
set -u
O=gpurun_out/r2c
mkdir -p $O
P="python tools/defaults_probe.py --docs 32768 --processes 16"
$P > $O/plain_defaults.json 2> $O/plain_defaults.err || { echo "plain run failed"; exit 1; }
KERN='regex:scvb0|phi_blend|hat_colsum|sum_partials|rho_table|colsum|perplexity|transpose|narrow|doc_wc|pcg_fill|inv_z|reciprocal'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERN" -c 600 --csv --log-file $O/launches_defaults.csv $P > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
LDA_B200_PAIR_MAX_DOCS=0 ncu --set full --clock-control none --import-source on -k regex:scvb0_docs32 -s 30 -c 1 -o $O/prof_unpaired $P > $O/ncu_unpaired.log 2>&1
echo "unpaired rc=$?"
LDA_B200_PAIR_MAX_DOCS=16384 ncu --set full --clock-control none --import-source on -k regex:scvb0_docs32 -s 30 -c 1 -o $O/prof_paired $P > $O/ncu_paired.log 2>&1
echo "paired rc=$?"
