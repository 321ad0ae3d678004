set -u
N="ncu --set full --clock-control none --import-source on -f"
for c in ${1:-5 44 64}; do
  $N -k regex:step_body_plane_ -s 1 -c 1 -o /tmp/sc$c python profiles/prof_strict_compact.py $c > /tmp/sc$c.log 2>&1
  ncu -i /tmp/sc$c.ncu-rep --page raw --csv > /tmp/sc$c.raw.csv 2>/dev/null && python profiles/ncu_extract.py /tmp/sc$c.raw.csv > gpurun_out/r2f_strict_compact_$c.csv
  tail -n 1 /tmp/sc$c.log
done
