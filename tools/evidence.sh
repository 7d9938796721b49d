#!/usr/bin/env bash
# One GPU-box call that refreshes the measured evidence under gpurun_out/ (copy what is to be
# judged into profiles/ afterwards).  Usage:
#   gpurun --timeout 1200 -- 'bash tools/evidence.sh r02a'
# Produces  <tag>_bench.json  <tag>_ref.json  <tag>_ba_launches.csv  <tag>_match.ncu-rep
#           <tag>_ba_kernels.ncu-rep  <tag>_chol_trace.txt
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --path ba --steps 3 --warmup 1 --no-cpu-baseline"

timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_ref.json 2>> $OUT/${TAG}_bench.err
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
XRB_TRACE=1 timeout 100 python tools/chol_check.py 2994,2993 3000,59 > $OUT/${TAG}_chol_trace.txt 2>&1
# launch list of one BA run (serialised, cold: compare shares, not absolutes)
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
    --log-file $OUT/${TAG}_ba_launches.csv $B > /dev/null 2>&1
# full captures: the Schur kernels and the factorisation kernels of one solve, the fused matcher
timeout 250 ncu --set full --clock-control none --import-source on \
    -k regex:"k_cam_blocks|k_lin|k_gather|k_backsub" -s 8 -c 6 -f -o $OUT/${TAG}_ba_kernels $B > /dev/null 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:"chol_update|chol_panel" -s 150 -c 8 \
    -f -o $OUT/${TAG}_chol $B > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_tc -c 1 -f -o $OUT/${TAG}_match \
    python bench.py --path match --match-images 48 --steps 4 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la $OUT | tail -12
