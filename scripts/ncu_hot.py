"""Hot regions of a kernel from an ncu --import-source report: SASS instructions grouped into runs, ranked by
executed-instruction count and stall samples.  usage: ncu_hot.py rep [kernel-id-index]"""
import csv, subprocess, sys
rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    try:
        ins.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["Warp Stall Sampling (All Samples)"]])))
    except (ValueError, IndexError):
        pass
tot_i = sum(i for _, i, _ in ins); tot_s = sum(s for _, _, s in ins)
print(f"{len(ins)} SASS instructions, {tot_i} executed, {tot_s} samples")
# segment into basic-block-like runs: boundaries where executed count changes by > 2x or at branches
W = int(sys.argv[3]) if len(sys.argv) > 3 else 40
segs = []
for a in range(0, len(ins), W):
    seg = ins[a:a + W]
    segs.append((a, sum(i for _, i, _ in seg), sum(s for _, _, s in seg)))
for a, i, s in sorted(segs, key=lambda t: -t[2])[:14]:
    ops = {}
    for src, ii, ss in ins[a:a + W]:
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        ops[op] = ops.get(op, 0) + ss
    top = ", ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda t: -t[1])[:6])
    print(f"  sass[{a:5d}..{a + W:5d}]  inst {100.0 * i / tot_i:5.1f}%  samples {100.0 * s / tot_s:5.1f}%   {top}")
