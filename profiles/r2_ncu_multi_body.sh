set -u
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
cap() { local name=$1 regex=$2 skip=$3; shift 3
    $N -k regex:$regex -s $skip -c 1 -o /tmp/$name "$@" > /tmp/$name.log 2>&1
    ncu -i /tmp/$name.ncu-rep --page raw --csv > /tmp/$name.raw.csv 2>/dev/null && python profiles/ncu_extract.py /tmp/$name.raw.csv > gpurun_out/$name.csv
    ncu -i /tmp/$name.ncu-rep --page source --csv > /tmp/$name.src.csv 2>/dev/null; python - /tmp/$name.src.csv gpurun_out/$name.source_top.txt <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
def col(n):
    for i, h in enumerate(hdr):
        if h.strip() == n: return i
    return None
ci, cs, cl = col("# Instructions Executed") or col("Instructions Executed"), col("Source"), col("#")
out = open(sys.argv[2], "w")
out.write("columns: " + " | ".join(hdr[:12]) + "\n")
if ci is not None and cs is not None:
    tot = 0; items = []
    for r in rows[1:]:
        try: n = float(r[ci])
        except Exception: continue
        tot += n; items.append((n, r[cs][:150]))
    items.sort(reverse=True)
    out.write(f"total warp instructions {tot:.3e}\n")
    for n, src in items[:40]:
        out.write(f"{100*n/max(tot,1):6.2f} %  {src}\n")
PY
    tail -n 1 /tmp/$name.log
}
cap r2f_multi_body step_multi_body_kernel 1 python profiles/prof_multi_body.py 32768 8
cap r2f_multi_body_b32 step_multi_body_kernel 1 python profiles/prof_multi_body.py 8192 32
ls -la gpurun_out/r2f_multi_body*
