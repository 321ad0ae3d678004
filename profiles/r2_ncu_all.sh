#!/bin/bash
# ncu --set full captures of the dominant kernel of every BASELINE config (round 2, final kernels).  Run on the GPU box:
#   bash profiles/r2_ncu_all.sh      (reports land in gpurun_out/, summaries are extracted with profiles/ncu_extract.py)
set -u
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:step_sphere_plane_pf_kernel -s 4 -c 1 -o gpurun_out/r2f_sphere_pf python bench.py --no-other-configs --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2f_sphere_pf.log 2>&1
$N -k regex:step_box_plane_pf_kernel -s 4 -c 1 -o gpurun_out/r2f_box_bounce python profiles/prof_cube.py bounce 128 > gpurun_out/r2f_box_bounce.log 2>&1
$N -k regex:step_box_plane_pf_kernel -s 4 -c 1 -o gpurun_out/r2f_box_incline python profiles/prof_cube.py incline 128 > gpurun_out/r2f_box_incline.log 2>&1
$N -k regex:step_two_ball_fast -s 2 -c 1 -o gpurun_out/r2f_two_ball python profiles/prof_two_ball.py > gpurun_out/r2f_two_ball.log 2>&1
$N -k regex:step_multi_sphere -s 0 -c 1 -o gpurun_out/r2f_ms_early python profiles/prof_multi_sphere.py 65536 0.0 > gpurun_out/r2f_ms_early.log 2>&1
$N -k regex:step_multi_sphere -s 4 -c 1 -o gpurun_out/r2f_ms_steady python profiles/prof_multi_sphere.py 65536 0.0 > gpurun_out/r2f_ms_steady.log 2>&1
$N -k regex:step_body_plane_kernel -s 1 -c 1 -o gpurun_out/r2f_strict_sphere python profiles/ab_strict.py > gpurun_out/r2f_strict_sphere.log 2>&1
tail -n 2 gpurun_out/r2f_*.log
