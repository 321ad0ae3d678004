"""Print the headline fields of bench.py JSON lines (one file per argument)."""
import json
import sys

for f in sys.argv[1:]:
    j = json.loads(open(f).read().strip().splitlines()[-1])
    r = j.get("roofline", {})
    print("%s: value %.4g  e2e %.4g  roofline %.3f (%s)  k1 %.3f" % (
        f, j["value"], j.get("e2e", {}).get("value", float("nan")), r.get("frac", float("nan")), r.get("kernel", "")[:48],
        j.get("roofline_k1", {}).get("frac", float("nan"))))
