set -u
N="ncu --set full --clock-control none --import-source on -f"
cap() { local name=$1 regex=$2 skip=$3; shift 3
    $N -k regex:$regex -s $skip -c 1 -o /tmp/$name "$@" > /tmp/$name.log 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > /tmp/$name.raw.csv 2>/dev/null && python profiles/ncu_extract.py /tmp/$name.raw.csv > gpurun_out/$name.csv
    tail -n 1 /tmp/$name.log
}
cap r2f_two_ball_ur step_two_ball_fast 2 python profiles/prof_two_ball.py
cap r2f_multi_body step_multi_body_kernel 1 python profiles/prof_multi_body.py 32768 8
cap r2f_multi_body_b32 step_multi_body_kernel 1 python profiles/prof_multi_body.py 8192 32
