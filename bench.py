#!/usr/bin/env python
"""Benchmark of the denoising hot path: DDIM-100 images/sec of Unet(dim=64, dim_mults=(1,2,4,8)) on 3x32x32 noise.

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path (torchrun for N > 1)
    python bench.py --impl reference ...                     # the CPU PyTorch restatement of the reference (oracle port)

One "step" = one complete 100-step DDIM sampling pass over one per-GPU batch of synthetic x_T (weak scaling: the
per-GPU batch is fixed, the global batch is N x that).  Prints ONE JSON line on rank 0.  See DESIGN.md section
"Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

DDIM_STEPS = 100
IMAGE = 32
CHANNELS = 3
MODEL_KW = dict(dim=64, dim_mults=(1, 2, 4, 8))
WORKLOAD = "DDIM-100 sampling, CIFAR Unet(dim=64, dim_mults=(1,2,4,8)), 3x32x32 (BASELINE configs[4] at image size 32)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        if os.environ.get("DDM_BENCH_NO_CLOCKS"):
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_model(seed=0):
    """Reference algorithm on CPU (oracle port) with the same synthetic weights as the CUDA arm."""
    from oracle import unet_forward, infer_config, synth_state_dict
    import diffusion_models_b200 as ddm
    shapes = {k: tuple(s) for k, (s, _) in ddm.Unet(**MODEL_KW).spec.params.items()}
    sd = synth_state_dict(shapes, seed)
    cfg = infer_config(sd)
    return lambda x, t, sc=None: unet_forward(sd, x, t, cfg)


def time_cpu_reference(batch, iters, warm=1):
    """Per-iteration time of the reference sampler's loop body (U-Net fp32 + DDIM update) on the host cores."""
    from oracle import make_schedule, ddim_time_pairs, ddim_update
    torch.set_num_threads(os.cpu_count())
    model, sch = oracle_model(), make_schedule(1000)
    x = torch.randn((batch, CHANNELS, IMAGE, IMAGE), generator=torch.Generator().manual_seed(1234))
    pairs = ddim_time_pairs(1000, DDIM_STEPS)
    times = []
    with torch.inference_mode():
        for i in range(warm + iters):
            t, tn = pairs[i]
            t0 = time.perf_counter()
            out = model(x, torch.full((batch,), t, dtype=torch.long))
            x, _ = ddim_update(sch, out, x, t, tn, 0.0, None)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 16
    per_step_iters = 2                       # a bounded sample of the 100-iteration workload per "step"
    all_t = time_cpu_reference(batch, per_step_iters * (args.steps + args.warmup), warm=0)
    timed = all_t[per_step_iters * args.warmup:]
    it = statistics.median(timed)
    value = batch / (DDIM_STEPS * it)
    sample = (f"B={batch}, {len(timed)} timed loop iterations (U-Net fp32 + DDIM update) of the 100 per image after "
              f"{per_step_iters * args.warmup} warm-up; images/s = B / (100 x median iteration time)")
    line = {"impl": "reference", "metric": "ddim100_images_per_sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": it * DDIM_STEPS * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": batch, "ddim_steps": DDIM_STEPS, "device": "host CPU"},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    import diffusion_models_b200 as ddm
    from diffusion_models_b200.flops import unet_flops_per_image
    from oracle import synth_state_dict                   # weights only: deterministic synthetic state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    shape = (B, CHANNELS, IMAGE, IMAGE)

    model = ddm.Unet(**MODEL_KW)
    model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0))
    model = model.to(dev).eval()
    diff = ddm.DenoisingDiffusion(model, image_size=IMAGE, sampling_timesteps=DDIM_STEPS).to(dev)

    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(shape, generator=gen).pin_memory()
    y_host = torch.empty((B * world,) + shape[1:]).pin_memory()
    x_dev = x_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def sample_resident():
        return ddm.sample_sharded(lambda b, r: diff.ddim_sample(shape, noise=x_dev), B * world)

    def sample_e2e():
        xd = x_host.to(dev, non_blocking=True)
        out = ddm.sample_sharded(lambda b, r: diff.ddim_sample(shape, noise=xd), B * world)
        y_host.copy_(out, non_blocking=True)
        return out

    def timed(fn, steps):
        total = 0.0
        for _ in range(steps):
            flush.zero_()                                            # L2 flush between timed iterations (not timed)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            total += e0.elapsed_time(e1)
        t = torch.tensor([total], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                 # max over ranks
        return t.item()

    for _ in range(args.warmup):
        sample_resident()
    torch.cuda.synchronize(dev)
    l0 = ddm._lib.launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(sample_resident, args.steps)
    launches_api = ddm._lib.launch_count() - l0
    graph_nodes = getattr(diff, "_last_graph_launches", 0)
    gpu_launches = launches_api + args.steps * graph_nodes
    sample_e2e()
    ms_e2e = timed(sample_e2e, args.steps)

    images = B * world * args.steps
    value = images / (ms / 1e3)
    e2e = images / (ms_e2e / 1e3)
    flops_img = unet_flops_per_image(model.spec, IMAGE, IMAGE) * DDIM_STEPS
    pk, pk_src = peaks()

    line = {"metric": "ddim100_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world, "ddim_steps": DDIM_STEPS,
                       "parallelism": f"batch-sharded x{world}, one final all-gather", "weights": "synthetic seed 0",
                       "l2": "256 MiB flush write between timed iterations; per-step working set >> 126 MB L2"},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": y_host.numel() * 4},
            "gpu_launches": int(gpu_launches), "clocks": clocks.summary()}

    if rank == 0:
        # roofline of the conv kernel family: algorithmic conv+linear FLOPs of the step / step time (whole step, so the
        # non-conv kernels count against it), plus the dominant layer shape timed alone.
        ach = value / world * flops_img / 1e12
        dom = time_dominant_conv(model, B, dev)
        line["roofline"] = {"bound": "tensor", "achieved": dom["tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": dom["tflops"] / pk["bf16_tflops"], "traffic": DOMINANT_DRAM_BYTES if B == 1024 else None,
                            "traffic_source": "dram__bytes_read+write of this launch in profiles/r01_ncu_full_conv_v10.txt (ncu --set full, B=1024)",
                            "algorithmic_bytes": 2 * B * IMAGE * IMAGE * 64 * 2, "peak_source": pk_src + " (burst, kernel timed alone)",
                            "kernel": "conv_tc_kernel", "layer": dom["layer"], "us_per_launch": dom["us"],
                            "whole_step": {"achieved": ach, "peak": pk["bf16_tflops_sustained"], "frac": ach / pk["bf16_tflops_sustained"],
                                           "unit": "TFLOP/s", "gflop_per_image": flops_img / 1e9,
                                           "note": "conv+linear algorithmic FLOPs / full sampling time, per GPU; peak = sustained"}}
        if world == 1:
            it = time_cpu_reference(16, 4, warm=1)
            med = statistics.median(it)
            line["cpu_baseline"] = {"value": 16 / (DDIM_STEPS * med), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "oracle port, B=16, 4 timed loop iterations after 1 warm-up; images/s = 16 / (100 x median)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ncu --set full capture of downs.0.0.block1 at B=1024 (profiles/r01_ncu_full_conv_v10.txt): 134.37 MB read + 90.12 MB written
# (the tail of the 134 MB output is still in L2 when the kernel ends)
DOMINANT_DRAM_BYTES = 134366208 + 90123264


def time_dominant_conv(model, B, dev, reps=20):
    """The FLOP-heaviest layer family (3x3, 64->64 @32x32 with the fused Block epilogue; 33.5 % of the network) alone."""
    eng = model.engine(B, IMAGE, IMAGE, time_rows=1, device=dev)
    tag = "downs.0.0.block1"
    op = dict(eng.ops)[tag]
    s = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(3):
        op(s)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op(s)
    e1.record()
    torch.cuda.synchronize(dev)
    us = e0.elapsed_time(e1) * 1e3 / reps
    flops = 2.0 * 64 * 64 * 9 * IMAGE * IMAGE * B
    return {"us": us, "tflops": flops / (us * 1e-6) / 1e12, "layer": f"{tag}: conv3x3 64->64 @32x32, B={B}, RMSNorm+scale/shift+SiLU epilogue"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("DDM_BENCH_BATCH", "1024")), help="per-GPU batch")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
