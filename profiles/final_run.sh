set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_final_ref.json 2> gpurun_out/bench_r1_final_ref.err; echo ref rc=$?
python bench.py --dtype fp32 --no-cpu-baseline > gpurun_out/bench_r1_final_fp32.json 2>/dev/null
python profiles/show_bench.py gpurun_out/bench_r1_final.json gpurun_out/bench_r1_final_fp32.json
python profiles/prof_two_ball.py; python profiles/prof_cube.py bounce; python profiles/prof_cube.py incline
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_final3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --k1-launches 8 > gpurun_out/ncu_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_sphere_plane_pf_kernel -s 24 -c 2 -f -o gpurun_out/prof_r1_final3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --k1-launches 8 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_sphere_plane_pf2 -s 24 -c 2 -f -o gpurun_out/prof_r1_final3_fp32 python bench.py --dtype fp32 --steps 2 --warmup 3 --no-cpu-baseline --k1-launches 8 > gpurun_out/ncu_f3_32.log 2>&1
python profiles/parity_report.py > gpurun_out/parity_report3.md 2> gpurun_out/parity_report3.err; echo parity rc=$?
python profiles/parity_divergence_fast.py > gpurun_out/parity_div_fast3.log 2>&1; echo div rc=$?
