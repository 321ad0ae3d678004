#!/bin/bash
# ncu --set full captures of the dominant kernel of every BASELINE config (round 2, final kernels).  Run on the GPU box:
#   bash profiles/r2_ncu_all.sh
# The reports stay in /tmp on the box (they exceed what gpurun copies back); the per-launch summaries extracted with
# profiles/ncu_extract.py land in gpurun_out/r2f_*.csv.
set -u
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
cap() {   # name, kernel regex, skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    $N -k regex:$regex -s $skip -c 1 -o /tmp/$name "$@" > /tmp/$name.log 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > /tmp/$name.raw.csv 2>/dev/null && python profiles/ncu_extract.py /tmp/$name.raw.csv > gpurun_out/$name.csv
    tail -n 1 /tmp/$name.log
}
cap r2f_sphere_pf step_sphere_plane_pf_kernel 4 python bench.py --no-other-configs --no-cpu-baseline --steps 1 --warmup 1
cap r2f_box_bounce step_box_plane_pf_kernel 4 python profiles/prof_cube.py bounce 128
cap r2f_box_incline step_box_plane_pf_kernel 4 python profiles/prof_cube.py incline 128
cap r2f_two_ball step_two_ball_fast 2 python profiles/prof_two_ball.py
cap r2f_ms_early step_multi_sphere 0 python profiles/prof_multi_sphere.py 65536 0.0
cap r2f_ms_steady step_multi_sphere 4 python profiles/prof_multi_sphere.py 65536 0.0
cap r2f_ms_early_mu03 step_multi_sphere 0 python profiles/prof_multi_sphere.py 65536 0.3
cap r2f_strict_sphere step_body_plane_kernel 1 python profiles/ab_strict.py
ls -la gpurun_out/r2f_*.csv
