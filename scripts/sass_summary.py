"""Per-kernel census of the SASS mnemonics that identify the Blackwell-native paths (B200_PROFILING.md "What proves a
Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, HMMA = legacy mma.sync.
usage: python scripts/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "diffusion-models_b200/libddm_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "SYNCS", "MUFU", "LDGSTS", "BAR.SYNC", "BAR.ARV"]
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for p in pats:
            if op.startswith(p):
                counts[cur][p] += 1
print(f"SASS census of {lib} (sm_100a): instruction counts per kernel")
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{p:>8s}" for p in pats))
for k, c in counts.items():
    print(f"{k[:58]:58s} {c['_total']:6d} " + " ".join(f"{c[p]:8d}" for p in pats))
