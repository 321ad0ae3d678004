set -u
N="ncu --set full --clock-control none --import-source on -f"
cap() { local name=$1 regex=$2 skip=$3; shift 3
    $N -k regex:$regex -s $skip -c 1 -o /tmp/$name "$@" > /tmp/$name.log 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > /tmp/$name.raw.csv 2>/dev/null && python profiles/ncu_extract.py /tmp/$name.raw.csv > gpurun_out/$name.csv
    tail -n 1 /tmp/$name.log
}
cap r3b_strict_cube_bounce_hybrid step_body_plane_resident_kernel 2 python profiles/prof_cube.py bounce 64 strict
cap r3b_strict_cube_incline_hybrid step_body_plane_resident_kernel 2 python profiles/prof_cube.py incline 64 strict
export RBS_STRICT_COMPACT=-1
cap r3b_strict_cube_bounce_thread_per_env step_body_plane_kernel 2 python profiles/prof_cube.py bounce 64 strict
cap r3b_strict_cube_incline_thread_per_env step_body_plane_kernel 2 python profiles/prof_cube.py incline 64 strict
