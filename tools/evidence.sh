#!/usr/bin/env bash
# One GPU-box call that refreshes the measured evidence under gpurun_out/ (copy what is to be
# judged into profiles/ afterwards).  Usage:
#   gpurun --timeout 1200 -- 'bash tools/evidence.sh r02a'
# Produces  <tag>_bench.json  <tag>_ref.json  <tag>_ba_launches.csv  <tag>_chol_trace.txt and the ncu reports
#           <tag>_{chol,schur,fm,filter,match,pose}.ncu-rep
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --path ba --steps 3 --warmup 1 --no-cpu-baseline"

timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_gpu_tests.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_ref.json 2>> $OUT/${TAG}_bench.err
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
XRB_TRACE=1 timeout 200 python tools/chol_check.py > $OUT/${TAG}_chol_trace.txt 2>&1
# launch list of one BA run (serialised, cold: compare shares, not absolutes)
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_ba_launches.csv $B --no-c4 > /dev/null 2>&1
# full captures: the factorisation kernels of one solve, the Schur kernels, the verification and filter kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_tile_cholesky|k_tile_backsolve" -s 4 -c 2 \
    -f -o $OUT/${TAG}_chol $B --no-c4 > /dev/null 2>&1
timeout 250 ncu --set full --clock-control none --import-source on \
    -k regex:"k_cam_blocks|k_lin|k_gather|k_backsub" -s 8 -c 5 -f -o $OUT/${TAG}_schur $B --no-c4 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_fm_loransac" -c 1 -f -o $OUT/${TAG}_fm \
    python -m pytest tests/test_fm_gpu.py -m gpu -q -k equals > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_filter_points" -c 1 -f -o $OUT/${TAG}_filter \
    python -m pytest tests/test_ba_gpu.py -q -k "filter and seq" > /dev/null 2>&1
# the matcher kernel (148 pairs, one per SM) and the pose-refinement kernel (4 096 poses)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_tc -s 1 -c 1 -f -o $OUT/${TAG}_match \
    python tools/ncu_match.py 3 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_pose_refine -s 1 -c 1 -f -o $OUT/${TAG}_pose \
    python -c "import bench; bench.pose_refine_rate(0, with_cpu=False)" > /dev/null 2>&1
ls -la $OUT | tail -16
