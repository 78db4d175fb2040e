#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun, AFTER the same commands have exited 0 without ncu).
# Full-set captures of one warm launch of every kernel that matters; reports land in gpurun_out/ and are summarised
# with profiles/summarize.py into profiles/r2_*_full.txt.
set -x
NCU="ncu --set full --clock-control none --import-source on"
O=gpurun_out
R=/tmp/ncu_r2   # the reports stay on the box (full-set reports with sources exceed what gpurun copies back); summaries travel
mkdir -p $R
python profiles/run_c2_once.py 2 > $O/run_c2.log 2>&1 || exit 1
$NCU -k regex:"coherence_probe|tile_hist|bin_totals|tile_scan|bin_cursors|tile_scatter|eval_brick|eval_deferred|unpermute|eval_kernel" --launch-skip 10 -c 10 -f -o $R/r2_c2_exact python profiles/run_c2_once.py 2 > $O/ncu_c2_exact.log 2>&1
$NCU -k regex:eval_brick --launch-skip 1 -c 1 -f -o $R/r2_c2_fast python profiles/run_c2_once.py 2 fast > $O/ncu_c2_fast.log 2>&1
$NCU -k regex:eval_kernel --launch-skip 1 -c 1 -f -o $R/r2_c3_plain_blocked python profiles/run_c34_once.py c3 > $O/ncu_c3.log 2>&1
$NCU -k regex:eval_kernel --launch-skip 1 -c 1 -f -o $R/r2_c4_plain_blocked python profiles/run_c34_once.py c4 > $O/ncu_c4.log 2>&1
$NCU -k regex:eval_kernel --launch-skip 1 -c 1 -f -o $R/r2_c1_plain python profiles/run_c1_once.py > $O/ncu_c1.log 2>&1
$NCU -k regex:"fast_|line_solve" --launch-skip 3 -c 3 -f -o $R/r2_c3_prefilter_fast python profiles/run_prefilter_once.py c3 fast > $O/ncu_pf_fast.log 2>&1
$NCU -k regex:"line_solve" --launch-skip 2 -c 2 -f -o $R/r2_c3_prefilter_exact python profiles/run_prefilter_once.py c3 > $O/ncu_pf_exact.log 2>&1
$NCU -k regex:"pad1|line_solve" --launch-skip 4 -c 4 -f -o $R/r2_c2_prefilter python profiles/run_prefilter_once.py c2 > $O/ncu_pf_c2.log 2>&1
$NCU -k regex:march --launch-skip 2 -c 2 -f -o $R/r2_c5 python profiles/run_c5_once.py > $O/ncu_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > $O/ncu_bench.log 2>&1
for r in $R/*.ncu-rep; do b=$(basename $r .ncu-rep); python profiles/summarize.py $r > $O/${b}_full.txt; done
ls -la $R/*.ncu-rep $O/*_full.txt
