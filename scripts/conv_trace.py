"""Device-side timeline of the conv kernel's roles for one layer (CTA 0): run with DDM_CONV_DEBUG=128 (| other bits)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusion_models_b200 as ddm

B = int(os.environ.get("B", "1024"))
tag = os.environ.get("TAG", "downs.0.0.block1")
model = ddm.Unet(dim=64, dim_mults=(1, 2, 4, 8)).cuda().eval()
eng = model.engine(B, 32, 32, time_rows=1)
eng.run_body(); torch.cuda.synchronize()
buf = (C.c_longlong * (2 * 5120))()
eng.lib.ddm_debug_conv_trace(buf, 5120)          # drain whatever the warm-up recorded
op = dict(eng.ops)[tag]
op(torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
n = eng.lib.ddm_debug_conv_trace(buf, 5120)
ev = sorted(((buf[2 * i + 1], buf[2 * i] >> 48, (buf[2 * i] >> 32) & 0xFFFF, buf[2 * i] & 0xFFFFFFFF) for i in range(n)))
t0 = ev[0][0]
roles = {0: "producer", 1: "issuerA", 2: "issuerB", 3: "epiG0", 4: "epiG1"}
names = {0: {0: "empty ok", 1: "loads issued"}, 1: {0: "acc_empty ok", 1: "full ok", 2: "token ok", 3: "mma issued", 4: "committed"},
         2: {0: "acc_empty ok", 1: "full ok", 2: "token ok", 3: "mma issued", 4: "committed"},
         3: {0: "acc_full ok", 1: "bar acc", 2: "pass1 done", 3: "bar pre", 4: "pass2 done", 5: "bar post"},
         4: {0: "acc_full ok", 1: "bar acc", 2: "pass1 done", 3: "bar pre", 4: "pass2 done", 5: "bar post"}}
lo, hi = int(os.environ.get("FROM", "2000")), int(os.environ.get("TO", "12000"))
print(f"TRACE {tag}: {n} events; showing cycles {lo}..{hi} after the first event")
for t, r, e, idx in ev:
    dt = t - t0
    if lo <= dt <= hi:
        print(f"TRACE {dt:8d} {roles.get(r, r):9s} {names.get(r, {}).get(e, e):14s} #{idx}")

# interval statistics (median over the trace, skipping the first 10 events of each kind)
import statistics
def med(pairs):
    return statistics.median(pairs[10:]) if len(pairs) > 12 else float("nan")
by = {}
for t, r, e, idx in ev:
    by.setdefault((r, e), []).append((idx, t))
def interval(r, e0, e1, same_idx=True):
    a, b = dict(by.get((r, e0), [])), dict(by.get((r, e1), []))
    return [b[k] - a[k] for k in sorted(a) if k in b and b[k] >= a[k]]
for r in (1, 2):
    print(f"STAT {roles[r]}: full->token {med(interval(r,1,2))}  token->issued {med(interval(r,2,3))}  issued->committed {med(interval(r,3,4))}")
pe = [t for _, t in by.get((0, 0), [])]; pl = [t for _, t in by.get((0, 1), [])]
print(f"STAT producer: issue {med([b - a for a, b in zip(pe, pl)])}  wait-empty {med([a2 - b for b, a2 in zip(pl, pe[1:])])}  per-stage {med([b - a for a, b in zip(pe, pe[1:])])}")
for r in (3, 4):
    print(f"STAT {roles[r]}: accfull->bar {med(interval(r,0,1))} pass1 {med(interval(r,1,2))} barpre {med(interval(r,2,3))} pass2 {med(interval(r,3,4))} barpost {med(interval(r,4,5))}")
    af = [t for _, t in by.get((r, 0), [])]
    print(f"STAT {roles[r]}: per-tile period {med([b - a for a, b in zip(af, af[1:])])}")

ae = [t for _, t in by.get((1, 0), [])]
print(f"STAT issuerA per-tile period (acc_empty ok): {med([b - a for a, b in zip(ae, ae[1:])])}  tiles {len(ae)}  span {ev[-1][0] - ev[0][0]}")

# hand-over between the two issuers: previous stage's "mma issued" -> this stage's "token ok"
tok = {}; iss = {}
for r in (1, 2):
    for idx, t in by.get((r, 2), []): tok[idx] = t
    for idx, t in by.get((r, 3), []): iss[idx] = t
ho = [tok[g] - iss[g - 1] for g in sorted(tok) if g - 1 in iss]
it = [iss[g] - tok[g] for g in sorted(tok) if g in iss]
per = [tok[g + 1] - tok[g] for g in sorted(tok) if g + 1 in tok]
print(f"STAT handover (issued g-1 -> token ok g): {med(ho)}   issue (token ok -> issued): {med(it)}   stage period: {med(per)}")
