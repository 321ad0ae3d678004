set -u
N="ncu --set full --clock-control none --import-source on -f"
cap() { local name=$1 regex=$2 skip=$3; shift 3
    $N -k regex:$regex -s $skip -c 1 -o /tmp/$name "$@" > /tmp/$name.log 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > /tmp/$name.raw.csv 2>/dev/null && python profiles/ncu_extract.py /tmp/$name.raw.csv > gpurun_out/$name.csv
    tail -n 1 /tmp/$name.log
}
python profiles/prof_multi_sphere.py 8192 0.0 strict 32
cap r3c_strict_ms_early step_multi_sphere_kernel 0 python profiles/prof_multi_sphere.py 8192 0.0 strict 32
cap r3c_strict_ms_steady step_multi_sphere_kernel 4 python profiles/prof_multi_sphere.py 8192 0.0 strict 32
