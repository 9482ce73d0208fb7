"""Join an ncu launch list (gpu__time_duration.sum per launch) with the engine's op tags; print per-op and per-kind shares."""
import csv
import json
import sys
from collections import defaultdict

csv_path, tags_path = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rows = []
with open(csv_path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
tags = json.load(open(tags_path))
body, tops = tags["body_ops"], tags["time_ops"]
n = len(body)
ours = [(k, t) for k, t in rows if "ddm" in k or "conv_tc" in k or "kernel" in k]
last = ours[-n:]                      # the last repetition of the body
assert len(last) == n
per = [(tag, k.split("(")[0].split("::")[-1], ns) for tag, (k, ns) in zip(body, last)]
total = sum(ns for _, _, ns in per)
print(f"B={tags['B']} img={tags['img']}: {n} launches, {total/1e6:.3f} ms per forward (serialised, cold cache)")
kinds = defaultdict(float)
for tag, k, ns in per:
    kinds[k] += ns
for k, ns in sorted(kinds.items(), key=lambda x: -x[1]):
    print(f"  {k:34s} {ns/1e6:8.3f} ms  {100*ns/total:5.1f}%")
meta = tags.get("meta", {})
print("top ops:   (TF = 2*M*N*K_pad / time; GB/s = (in+out bytes) / time)")
for tag, k, ns in sorted(per, key=lambda x: -x[2])[:60]:
    extra = ""
    if tag in meta:
        m = meta[tag]
        tf = 2.0 * m["M"] * m["N"] * m["K"] / ns / 1e3
        gbs = (m["in_bytes"] + m["out_bytes"]) / ns
        extra = f" M={m['M']:8d} N={m['N']:4d} K={m['K']:5d} {tf:7.1f} TF {gbs:7.0f} GB/s"
    print(f"  {tag:30s} {k:22s} {ns/1e3:8.1f} us {100*ns/total:5.1f}%{extra}")
