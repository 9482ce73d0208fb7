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


def cpu_ddim100(batch):
    """One complete DDIM-100 sampling call of the reference algorithm (oracle port) on the host cores; returns seconds."""
    from oracle import make_schedule, ddim_sample
    torch.set_num_threads(os.cpu_count())
    model, sch = oracle_model(), make_schedule(1000)
    x = torch.randn((batch, CHANNELS, IMAGE, IMAGE), generator=torch.Generator().manual_seed(1234))
    t0 = time.perf_counter()
    with torch.inference_mode():
        ddim_sample(model, sch, x, DDIM_STEPS)
    return time.perf_counter() - t0


def run_reference(args):
    """The reference's CPU implementation of the path (the oracle port: the reference is Python and cannot travel to the GPU
    box), on all host cores.  One step = ONE COMPLETE DDIM-100 sampling call, timed as such (no
    extrapolation); the batch is the bounded sample: small enough that `--steps K --warmup W` calls end within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = int(os.environ.get("DDM_REF_BATCH", "0"))
    if batch <= 0:       # calibrate: the largest batch of {1, 2, 4} whose DDIM-100 call stays under ~8 s on this host
        it = statistics.median(time_cpu_reference(4, 2, warm=1))
        batch = 4 if it * DDIM_STEPS <= 8.0 else (2 if it * DDIM_STEPS <= 14.0 else 1)
    for _ in range(args.warmup):
        cpu_ddim100(batch)
    times = [cpu_ddim100(batch) for _ in range(args.steps)]
    total = sum(times)
    value = batch * args.steps / total
    sample = (f"B={batch} per call, {args.steps} complete DDIM-100 sampling calls (100 fp32 U-Net evaluations + DDIM updates each) timed "
              f"after {args.warmup} warm-up calls; images/s = B x calls / total time (the CPU's best rate, at B=16, is in the GPU "
              f"arm's cpu_baseline)")
    line = {"impl": "reference", "metric": "ddim100_images_per_sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": batch, "global_batch": batch, "ddim_steps": DDIM_STEPS,
                       "device": "host CPU", "weights": "synthetic seed 0"},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def time_gpu_stock_pytorch(dev, batch, iters=4, warm=2):
    """Second comparator: the same algorithm (oracle port) through stock PyTorch
    on this B200 -- cuDNN / cuBLAS library kernels -- in fp32 with TF32 off (the reference's sampling precision) and under
    bf16 autocast.  Bounded: `iters` loop iterations (U-Net + DDIM update) each, extrapolated to the 100 of a sampling call."""
    from oracle import unet_forward, infer_config, synth_state_dict, make_schedule, ddim_time_pairs, ddim_update
    import diffusion_models_b200 as ddm
    shapes = {k: tuple(s) for k, (s, _) in ddm.Unet(**MODEL_KW).spec.params.items()}
    sd = {k: v.to(dev) for k, v in synth_state_dict(shapes, 0).items()}
    cfg, sch = infer_config(sd), make_schedule(1000)
    sch = type(sch)(**{f: getattr(sch, f).to(dev) for f in sch.__dataclass_fields__})
    pairs = ddim_time_pairs(1000, DDIM_STEPS)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    out = {}
    try:
        for name, autocast in (("fp32_tf32_off", False), ("bf16_autocast", True)):
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.benchmark = True
            x = torch.randn((batch, CHANNELS, IMAGE, IMAGE), generator=torch.Generator().manual_seed(1234)).to(dev)
            ts = []
            with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                for i in range(warm + iters):
                    t, tn = pairs[i]
                    torch.cuda.synchronize(dev)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    eps = unet_forward(sd, x, torch.full((batch,), t, dtype=torch.long, device=dev), cfg).float()
                    x, _ = ddim_update(sch, eps, x, t, tn, 0.0, None)
                    e1.record()
                    torch.cuda.synchronize(dev)
                    if i >= warm:
                        ts.append(e0.elapsed_time(e1) * 1e-3)
            out[name] = batch / (DDIM_STEPS * statistics.median(ts))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    return {"unit": "images/s", "fp32_tf32_off": out["fp32_tf32_off"], "bf16_autocast": out["bf16_autocast"], "kind": "port",
            "sample": f"oracle port on cuda through stock PyTorch {torch.__version__} (cuDNN/cuBLAS), B={batch}, {iters} timed loop "
                      f"iterations after {warm} warm-up per precision; images/s = B / (100 x median iteration time)"}


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
    scaling = "weak"
    if args.global_batch:                                   # strong scaling: the global batch is fixed, split over the ranks
        assert args.global_batch % world == 0, "--global-batch must divide by the number of GPUs"
        B, scaling = args.global_batch // world, "strong"
    shape = (B, CHANNELS, IMAGE, IMAGE)

    model = ddm.Unet(**MODEL_KW)
    model.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=0))
    model = model.to(dev).eval()
    diff = ddm.DenoisingDiffusion(model, image_size=IMAGE, sampling_timesteps=DDIM_STEPS).to(dev)

    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(shape, generator=gen).pin_memory()
    y_host = torch.empty(shape).pin_memory()                 # every rank reads back its own shard of the gathered result
    x_dev = x_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def sample_resident():
        return ddm.sample_sharded(lambda b, r: diff.ddim_sample(shape, noise=x_dev), B * world)

    def sample_e2e():
        xd = x_host.to(dev, non_blocking=True)
        out = ddm.sample_sharded(lambda b, r: diff.ddim_sample(shape, noise=xd), B * world)
        y_host.copy_(out[rank * B:(rank + 1) * B], non_blocking=True)
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
    gpu_launches = (launches_api + args.steps * graph_nodes) * world          # whole job: every rank launches the same list
    sample_e2e()
    ms_e2e = timed(sample_e2e, args.steps)

    images = B * world * args.steps
    value = images / (ms / 1e3)
    e2e = images / (ms_e2e / 1e3)
    flops_img = unet_flops_per_image(model.spec, IMAGE, IMAGE) * DDIM_STEPS
    pk, pk_src = peaks()

    line = {"metric": "ddim100_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world, "ddim_steps": DDIM_STEPS,
                       "parallelism": f"batch-sharded x{world}, one final all-gather", "weights": "synthetic seed 0",
                       "l2": "256 MiB flush write between timed iterations; per-step working set >> 126 MB L2"},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4 * world,
                    "d2h_bytes_per_step": y_host.numel() * 4 * world},
            "gpu_launches": int(gpu_launches), "clocks": clocks.summary()}

    if rank == 0:
        # roofline of the conv kernel family: algorithmic conv+linear FLOPs of the step / step time (whole step, so the
        # non-conv kernels count against it), plus the dominant layer shape timed alone.
        ach = value / world * flops_img / 1e12
        dom = time_dominant_conv(model, B, dev)
        traffic, traffic_src = dominant_traffic(B)
        line["roofline"] = {"bound": "tensor", "achieved": dom["tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": dom["tflops"] / pk["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_bytes": 2 * B * IMAGE * IMAGE * 64 * 2, "peak_source": pk_src + " (burst, kernel timed alone)",
                            "kernel": "conv_tc_kernel", "layer": dom["layer"], "us_per_launch": dom["us"],
                            "whole_step": {"achieved": ach, "peak": pk["bf16_tflops_sustained"], "frac": ach / pk["bf16_tflops_sustained"],
                                           "unit": "TFLOP/s", "gflop_per_image": flops_img / 1e9,
                                           "note": "conv+linear algorithmic FLOPs / full sampling time, per GPU; peak = sustained"}}
        if world == 1:
            it = time_cpu_reference(16, 30, warm=2)
            med = statistics.median(it)
            line["cpu_baseline"] = {"value": 16 / (DDIM_STEPS * med), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "oracle port (reference algorithm, fp32, all host cores), B=16, 30 timed loop iterations "
                                              "(U-Net + DDIM update) after 2 warm-up; images/s = 16 / (100 x median iteration time)"}
            if not os.environ.get("DDM_BENCH_NO_GPU_BASELINE"):
                del diff, model
                torch.cuda.empty_cache()
                line["gpu_baseline"] = time_gpu_stock_pytorch(dev, min(B, 1024))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dominant_traffic(B):
    """DRAM bytes of the dominant launch, from the `ncu --set full` capture of the CURRENT kernels that scripts/ncu_traffic.py
    distilled into profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch).  null when the capture
    is missing or was taken at another batch."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            rec = json.load(f)["downs.0.0.block1"]
        if rec.get("batch") != B:
            return None, f"profiles/traffic.json holds B={rec.get('batch')}, this run is B={B}"
        return int(rec["dram_read"] + rec["dram_write"]), rec["source"]
    except Exception:
        return None, "no ncu capture of the current build in profiles/traffic.json"


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
    ap.add_argument("--batch", type=int, default=int(os.environ.get("DDM_BENCH_BATCH", "1024")), help="per-GPU batch (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: fixed global batch split over the GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
