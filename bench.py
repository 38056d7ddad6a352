#!/usr/bin/env python
"""Benchmark of the hot path: samples/sec of with_logabsdet_jacobian through an 8-layer Planar flow, D=128,
N=2^20 Float32 per GPU (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input.  Prints ONE JSON line (rank 0).
  value      whole-job samples/s with the batch resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public API with HOST (pinned) buffers: H2D + kernels + D2H per step
  roofline   dominant kernel (the fused chain kernel): algorithmic bytes per launch / measured launch time
  cpu_baseline  the C restatement of the reference CPU path (oracle/b2b_oracle.c) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

D, NCOLS, NLAYERS = 128, 1 << 20, 8
METRIC = "samples/sec: with_logabsdet_jacobian through 8-layer Planar flow, D=128"
WORKLOAD = "Composed(8x PlanarLayer), D=128, N=2^20 per GPU, Float32 (BASELINE configs[1])"
CPU_SAMPLE_COLS = 1 << 18  # per step of the --impl reference arm


def planar_params(seed_base=100):
    """per layer w,u ~ N(0,1)/sqrt(D), b ~ N(0,1), PCG64 seeds 100+layer."""
    out = []
    for l in range(NLAYERS):
        rng = np.random.Generator(np.random.PCG64(seed_base + l))
        w = (rng.standard_normal(D) / np.sqrt(D)).astype(np.float32)
        u = (rng.standard_normal(D) / np.sqrt(D)).astype(np.float32)
        b = rng.standard_normal(1).astype(np.float32)
        out.append((w, u, b))
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def pick_threads(layers):
    """Host threads for the CPU restatement: the fastest of {all cores the process may use, 64, 32, 16, 8} on a small
    probe (with fresh outputs per layer the pass is page-fault / bandwidth bound, so more threads is not always faster)."""
    from oracle import oracle_c as C

    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = C.num_procs()
    avail = max(1, min(avail, C.num_procs()))
    rng = np.random.Generator(np.random.PCG64(2))
    probe = np.asfortranarray(rng.standard_normal((D, 1 << 15), dtype=np.float32))
    best_t, best = avail, float("inf")
    for t in sorted({avail, *[c for c in (64, 32, 16, 8) if c <= avail]}, reverse=True):
        C.planar_chain_fwd(layers, probe, nthreads=t)
        t0 = time.perf_counter()
        C.planar_chain_fwd(layers, probe, nthreads=t)
        dt = time.perf_counter() - t0
        if dt < best:
            best_t, best = t, dt
    return best_t


def cpu_baseline(nthreads=None, repeats=5, cols=NCOLS):
    """Times the C restatement of the reference CPU path on a bounded sample of the same workload."""
    from oracle import oracle_c as C

    layers = planar_params()
    nthreads = nthreads or pick_threads(layers)
    rng = np.random.Generator(np.random.PCG64(1))
    x = np.asfortranarray(rng.standard_normal((D, cols), dtype=np.float32))
    C.planar_chain_fwd(layers, x[:, : 1 << 12], nthreads=nthreads)  # warm-up (library load, threads)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        C.planar_chain_fwd(layers, x, nthreads=nthreads)
        best = min(best, time.perf_counter() - t0)
    return {
        "value": cols / best,
        "unit": "samples/s",
        "cores": int(nthreads),
        "kind": "port",
        "sample": f"{cols} of {NCOLS} columns, best of {repeats} passes; C/OpenMP restatement of the reference "
                  f"CPU pass structure (gemv pass + broadcast pass per layer), not the Julia package; thread count "
                  f"picked by a probe",
        "seconds_per_pass": best,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Julia and no
    Julia toolchain exists in this image, so this arm times the C restatement of its CPU path (oracle port)
    with all host threads.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_c as C

    layers = planar_params()
    nthreads = pick_threads(layers)
    cols = CPU_SAMPLE_COLS
    rng = np.random.Generator(np.random.PCG64(1))
    x = np.asfortranarray(rng.standard_normal((D, cols), dtype=np.float32))
    for _ in range(max(args.warmup, 1)):
        C.planar_chain_fwd(layers, x[:, : 1 << 14], nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        C.planar_chain_fwd(layers, x, nthreads=nthreads)
    dt = time.perf_counter() - t0
    val = cols * args.steps / dt
    sample = f"{cols} of {NCOLS} columns per step; C/OpenMP restatement of the reference CPU path (oracle port)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": int(nthreads), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (0 auto, 1 v0 lane-group, 2 v1 interpreter, 3 unrolled planar)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default min(steps, 10))")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile", action="store_true",
                    help="profiling runs (under ncu): only the warm-up and the timed headline steps, no JSON contract line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import bijectors_jl_b200 as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, (world, args.gpus)
    B.lib().b2b_set_kernel_variant(args.variant)
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- synthetic workload (identical on every rank: weak scaling, N columns per GPU) ------------------
    flow = B.Composed(*[B.PlanarLayer(w, u, b) for (w, u, b) in planar_params()])
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.randn((NCOLS, D), device="cuda", generator=gen).t()  # D x N column-major, 512 MiB > L2
    y = B.colmajor_empty(D, NCOLS)
    lj = torch.empty(NCOLS, device="cuda")

    def step():
        B.run_chain(flow, x, y=y, logjac=lj)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    launches_per_step = B.lib().b2b_last_launch_count()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    clocks = sampler.stop() if sampler else None
    value = world * NCOLS * steps / (ms_total * 1e-3)

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_total / steps, "value": value}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- same chain with HOST-resident parameters (the reference's own residency): b2b_planar_chain_hostparams_f32 -
    host_flow = B.Composed(*[lay.to("cpu") for lay in B.flatten(flow)])
    for _ in range(3):
        B.run_chain(host_flow, x, y=y, logjac=lj)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        B.run_chain(host_flow, x, y=y, logjac=lj)
    e1.record()
    torch.cuda.synchronize()
    ms_hostparams = e0.elapsed_time(e1) / steps

    # ---- per-layer launches (the reference's launch structure: 8 kernels, y and logjac round-trip HBM) ---
    layers = B.flatten(flow)

    def step_layerwise():
        B.run_chain(layers[0], x, y=y, logjac=lj)
        for lay in layers[1:]:
            B.run_chain(lay, y, y=y, logjac=lj, accumulate=True)

    for _ in range(2):
        step_layerwise()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step_layerwise()
    e1.record()
    torch.cuda.synchronize()
    ms_layerwise = e0.elapsed_time(e1) / steps

    # ---- e2e: the public API with HOST buffers (pinned), H2D + kernels + D2H inside the timed region -----
    e2e_steps = args.e2e_steps or min(steps, 10)
    xh = torch.empty((NCOLS, D), dtype=torch.float32, pin_memory=True).t()
    xh.copy_(x)
    for _ in range(2):
        yh, ljh = B.with_logabsdet_jacobian(flow, xh)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        yh, ljh = B.with_logabsdet_jacobian(flow, xh)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * NCOLS * e2e_steps / float(dt)
    assert torch.equal(ljh.cuda(), lj), "host-buffer path and device path disagree"

    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_step = ms_total / steps
        bytes_fused = NCOLS * 4 * (2 * D + 1)  # read column + write column + write logjac, once per chain launch
        achieved = bytes_fused / (ms_step * 1e-3) / 1e9
        bytes_layerwise = NCOLS * 4 * (NLAYERS * (2 * D + 1) + (NLAYERS - 1))
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("chain_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "chain": "one fused chain launch per step (column read once, written once); parameters device-resident",
                       "l2": "inputs_larger_than_L2 (x, y 512 MiB each per GPU; L2 126 MB)",
                       "parallelism": f"columns sharded, {world} rank(s), no data-path collective",
                       "kernel_variant": args.variant},
            "gpu_launches": int(launches_per_step * steps),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(NCOLS * D * 4),
                    "d2h_bytes_per_step": int(NCOLS * D * 4 + NCOLS * 4), "steps": e2e_steps,
                    "api": "with_logabsdet_jacobian(flow, pinned host D x N) -> b2b_chain_run_host_f32"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
                         "kernel": "planar_dev_kernel: fused 8-layer chain, 1 launch/step", "algorithmic_bytes_per_launch": bytes_fused,
                         "accounting": "chain-fused: 4*(2D+1) B/sample per launch"},
            "host_resident_parameters": {"ms_per_step": ms_hostparams, "samples_per_s": NCOLS / (ms_hostparams * 1e-3),
                                         "frac": NCOLS * 4 * (2 * D + 1) / (ms_hostparams * 1e-3) / 1e9 / peak,
                                         "api": "b2b_planar_chain_hostparams_f32 (parameters as kernel arguments)"},
            "per_layer_launches": {"ms_per_step": ms_layerwise, "samples_per_s": NCOLS / (ms_layerwise * 1e-3),
                                   "achieved_gbs": bytes_layerwise / (ms_layerwise * 1e-3) / 1e9,
                                   "frac": bytes_layerwise / (ms_layerwise * 1e-3) / 1e9 / peak,
                                   "accounting": "8 launches: 4*(L*(2D+1)+(L-1)) B/sample"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
