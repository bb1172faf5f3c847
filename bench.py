#!/usr/bin/env python
"""Throughput of the fused thin-film D2Q9 LBM step on B200 -- the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--L 8192] [--workload film|thermal]

One "step" = one pass of the hot path (src/simulate.jl:15-22) over the whole lattice.  Default workload: the
8192 x 8192 thin film the north_star target is quoted on (tau = 1, Taumucs defaults, flat film + perturbation,
SURVEY.md 8d C5/C4 geometry); at N > 1 every rank owns an L x L slab (weak scaling, global lattice L x N*L) and the
ranks exchange halo rows over NCCL.  Prints ONE JSON line (rank 0).

  value     MLUPS with the state resident in HBM, CUDA-event timed, max over ranks
  e2e       MLUPS of a whole user-level job through the public API: pinned-host initial height -> H2D ->
            time_loop (K steps, fused kernels, per-tdump mass read-back) -> D2H of the final height
  roofline  144 B/LU (9 populations read + 9 written) x L^2 per launch / mean kernel time, against
            the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the C oracle (restated reference CPU path; Julia is not installable here) on the host cores,
            on a bounded 2048^2 sample of the same workload
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = 144.0  # algorithmic bytes per lattice update


def initial_height(L, Ly=None, j0=0, Ly_global=None, workload="film"):
    """Synthetic initial conditions of SURVEY.md 8d as NumPy F-arrays (rows j0 .. j0+Ly of a L x Ly_global lattice).
    film/thermal: h = 1 + 1e-3 sin(2π i/Lx) sin(2π j/Ly); droplet: spherical cap (singledroplet, radius min(L/4, 256), θ0 = 1/6,
    precursor 0.05); spinodal: h = 1 + 0.01 N(0,1), seed 20261017."""
    import numpy as np

    Ly = Ly or L
    Ly_global = Ly_global or Ly
    i = np.arange(L, dtype=np.float64)[:, None]
    j = (j0 + np.arange(Ly, dtype=np.float64))[None, :]
    if workload == "droplet":
        radius, c = min(L / 4.0, 256.0), (L // 2, Ly_global // 2)  # C2: radius 256 at 1024^2 (larger caps are unstable)
        circ = np.sqrt((i + 1 - c[0]) ** 2 + (j + 1 - c[1]) ** 2)
        inside = circ <= radius
        cap = (np.cos(np.arcsin(np.where(inside, circ / radius, 0.0))) - math.cos(math.pi / 6)) * radius
        h = np.where(inside, cap, 0.05)
        return np.asfortranarray(np.where(h < 0, 0.05, h))
    if workload == "spinodal":
        rng = np.random.default_rng([20261017, j0])
        return np.asfortranarray(1.0 + 0.01 * rng.standard_normal((L, Ly)))
    amp = 0.1 if workload == "thermal_moving" else 1e-3
    return np.asfortranarray(1.0 + amp * np.sin(2 * np.pi * i / L) * np.sin(2 * np.pi * j / Ly_global))


TMOVE = 98  # C4: the contact-angle pattern moves by (1, 1) every 98 steps (scripts/Moving_wettability_structs.jl:182-200)


def theta_pattern(L, Ly=None, j0=0, Ly_global=None):
    """C4 contact-angle field θ[i,j] = 1/9 + (1/36) sin(2π 2(i-1)/Lx) sin(2π 2(j-1)/Ly) (rows j0 .. j0+Ly)."""
    import numpy as np

    Ly = Ly or L
    Ly_global = Ly_global or Ly
    i = np.arange(L, dtype=np.float64)[:, None]
    j = (j0 + np.arange(Ly, dtype=np.float64))[None, :]
    return np.asfortranarray(1 / 9 + (1 / 36) * np.sin(2 * np.pi * 2 * i / L) * np.sin(2 * np.pi * 2 * j / Ly_global))


def segments(s0, n, period=TMOVE):
    """Split steps [s0, s0+n) at the multiples of `period`: [(first step, count, move the substrate afterwards?)]."""
    out, t, end = [], s0, s0 + n
    while t < end:
        nxt = min(end, (t // period + 1) * period)
        out.append((t, nxt - t, nxt % period == 0))
        t = nxt
    return out


def workload_params(args, K):
    """Taumucs keyword arguments of the workload."""
    kw = dict(Tmax=K, tdump=max(1, K // 2))
    if args.workload == "thermal":
        kw.update(kbt=1e-7)
    elif args.workload == "thermal_moving":
        kw.update(kbt=1e-7, γ=0.01, δ=1.0, μ=1 / 12, n=3, m=2, hmin=0.07)
    elif args.workload == "droplet":
        kw.update(n=3, m=2, hmin=0.07, δ=1.0)
    elif args.workload == "spinodal":
        kw.update(n=3, m=2, hmin=0.07, γ=0.01)
    return kw


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 50 ms by a background process that is started BEFORE the
    warm-up (nvidia-smi takes a while to come up); summary() keeps the samples that fall inside the timed window."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, enabled=True):
        self.rows, self.proc, self.index, self.t0, self.t1 = [], None, index, None, None
        if not enabled:  # ranks > 0: same interface, no nvidia-smi process
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __enter__(self):  # the timed window
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.proc = None

    def summary(self):
        self.stop()
        ok = [(t, r) for t, r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        inside = [r for t, r in ok if self.t0 is not None and self.t0 <= t <= self.t1 + 0.06]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: take the sample closest to it
            mid = 0.5 * (self.t0 + self.t1)
            inside = [min(ok, key=lambda tr: abs(tr[0] - mid))[1]]
            window = "nearest sample"
        sm = sorted(float(r[0]) for r in inside)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in inside)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(inside[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in inside), "samples": len(sm), "window": window}


def cpu_baseline_run(L, steps, threads, warmup=1, workload="film"):
    """The C oracle (faithful pass structure of the reference's CPU path) on an L x L sample; returns MLUPS.
    (thermal: the deterministic part only -- Julia's randn! is not restated.)"""
    from oracle import oracle_c as oc
    from oracle import oracle_np as onp

    oc.build()
    st = onp.State(L, L)
    st.height[...] = initial_height(L, workload=workload)
    okw = {"droplet": dict(n=3, m=2, hmin=0.07), "spinodal": dict(n=3, m=2, hmin=0.07, gamma=0.01),
           "thermal_moving": dict(n=3, m=2, hmin=0.07, gamma=0.01, mu=1 / 12)}.get(workload, {})
    p = onp.Params(**okw)
    lkw = dict(threads=threads)
    if workload == "thermal_moving":  # the contact-angle field (cospi evaluated once on the host, as the oracle takes it)
        import numpy as np

        lkw["cospi_theta"] = np.asfortranarray(np.cos(np.pi * theta_pattern(L)))
    oc.time_loop(st, p, nsteps=warmup, **lkw)
    t0 = time.perf_counter()
    oc.time_loop(st, p, nsteps=steps, **lkw)
    dt = time.perf_counter() - t0
    return L * L * steps / dt / 1e6, dt


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (restated in C: Julia cannot be installed offline) timed on
    this box's host cores with every thread OpenMP gives it, each step = one LBM step over a bounded 2048^2 sample."""
    if rank != 0:
        return
    from oracle import oracle_c as oc

    threads = oc.max_threads()
    Ls = 2048
    mlups, dt = cpu_baseline_run(Ls, args.steps, threads, warmup=max(1, min(args.warmup, 3)), workload=args.workload)
    m1, _ = cpu_baseline_run(1024, 3, 1, workload=args.workload)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(mlups, 3), "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": round(mlups, 3), "unit": "MLUPS", "cores": threads, "kind": "port",
                         "sample": f"{Ls}x{Ls} sample of the workload, {args.steps} steps, OpenMP x{threads}; "
                                   f"single-thread (Julia-like) rate on 1024^2: {m1:.2f} MLUPS"},
        "e2e": {"value": round(mlups, 3), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    what = {"film": "Taumucs defaults (n=9,m=3,theta=1/9), flat film + sine perturbation (SURVEY 8d C5)",
            "thermal": "Taumucs defaults + thermal noise kbt=1e-7 generated in-kernel (Philox), flat film + sine (C4 noise only)",
            "thermal_moving": "C4 complete: thermal noise kbt=1e-7 (Philox, in-kernel) + contact-angle pattern theta(x,y) moved "
                              "by (1,1) every 98 steps on the device, n=3,m=2,hmin=0.07,gamma=0.01,mu=1/12, h=1+0.1 sin sin",
            "droplet": "n=3,m=2,hmin=0.07,theta=1/9, spherical-cap droplet radius min(L/4,256) on a 0.05 precursor (C2)",
            "spinodal": "n=3,m=2,hmin=0.07,gamma=0.01, randomly perturbed film h=1+0.01 N(0,1) (C3)"}[args.workload]
    weak = args.gpus == 1 or getattr(args, "scaling", "weak") == "weak"
    rows = args.L if weak else args.L // args.gpus
    return {"workload": f"thin-film D2Q9 LBM step, {args.L}x{rows} per GPU, tau=1, {what}",
            "grid": [args.L, rows * args.gpus], "per_gpu_grid": [args.L, rows], "bytes_per_update_alg": B_ALG,
            "decomposition": "row slabs along y, NCCL send/recv halos" if args.gpus > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (>= 1.5 GB touched per step vs 126 MB L2)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--L", type=int, default=8192)
    ap.add_argument("--workload", default="film", choices=["film", "thermal", "thermal_moving", "droplet", "spinodal"],
                    help="film: C5 flat film + sine (default); thermal: film + Philox noise; thermal_moving: the complete C4 "
                         "(noise + contact-angle pattern moved by (1,1) every 98 steps, n=3,m=2,hmin=0.07,mu=1/12); droplet: C2 spherical cap, "
                         "n=3,m=2,hmin=0.07 (use --L 1024); spinodal: C3 random film, n=3,m=2,hmin=0.07,gamma=0.01 (--L 4096)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = L x L per rank (default, the contract), strong = L x L in total (L/N rows per rank)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch

    import __graft_entry__ as ge

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback on the b200 arm)")
    torch.cuda.set_device(local_rank)
    # nvidia-smi needs seconds to come up on a busy 8-GPU box: start the clock sampler now, on rank 0 only
    early_sampler = ClockSampler(local_rank, enabled=rank == 0)
    if rank == 0:
        ge.build()
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import swalbe_b200 as sw
    from swalbe_b200 import _lib

    lib = _lib.load()
    L, K, W = args.L, args.steps, args.warmup
    prm = sw.Taumucs(**workload_params(args, K))
    sysc = sw.SysConst(Lx=L, Ly=L, param=prm)
    rows = L if (world == 1 or args.scaling == "weak") else L // world   # rows per rank
    Ly_glob = rows * world
    thermal_seed = 1234 if args.workload in ("thermal", "thermal_moving") else None
    moving = args.workload == "thermal_moving"
    e2e = None

    if world == 1:
        st = sw.Sys(sysc, "GPU", kind="thermal" if thermal_seed is not None else "simple")
        h0 = initial_height(L, workload=args.workload)
        st.height.set(h0)
        th = inp = None
        if moving:
            th, inp = sw.Field(L, L).set(theta_pattern(L)), sw.Field(L, L)
            inp.set(th)

        def run(n, s0=0, **kw):
            if not moving:
                return sw.fused_steps(st, sysc, n, thermal_seed=thermal_seed, step0=s0,
                                      pressure_variant=_lib.PRESSURE_POWER_BROAD, **kw)
            segs = segments(s0, n)
            for q, (t0, cnt, move) in enumerate(segs):  # like the drivers' chunks: intermediate fields on the last one only
                sw.fused_steps(st, sysc, cnt, thermal_seed=thermal_seed, step0=t0, θ=th, skip_aux=q < len(segs) - 1,
                               pressure_variant=_lib.PRESSURE_POWER_BROAD, **kw)
                if move:
                    sw.move_substrate(th, inp, t0 + cnt, TMOVE)
        sampler = early_sampler
        run(W)
        torch.cuda.synchronize()
        l0 = lib.swalbe_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler as clocks:
            e0.record()
            run(K, W)
            e1.record()
            torch.cuda.synchronize()
        launches = int(lib.swalbe_launch_count() - l0)
        sampler.stop()
        ms = e0.elapsed_time(e1)
        mass_drift = abs(st.height.t.sum().item() - h0.sum()) / h0.sum()
        lu = L * L * K
        # moments-only row (populations materialised on the last step only), reported separately
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        run(W, lazy_populations=True)
        e2.record()
        run(K, W, lazy_populations=True)
        e3.record()
        torch.cuda.synchronize()
        lazy_mlups = lu / (e2.elapsed_time(e3) * 1e-3) / 1e6
        if not args.no_e2e:
            # e2e: the call a user makes -- host initial condition in, time_loop, host result out
            h_host = torch.from_numpy(np.ascontiguousarray(h0.transpose())).pin_memory()
            out_host = torch.empty_like(h_host).pin_memory()

            def job():
                st.height.t.copy_(h_host, non_blocking=True)
                st.velx.t.zero_(); st.vely.t.zero_()
                sw.equilibrium(st, sysc)
                if thermal_seed is None:
                    sw.time_loop(sysc, st)
                else:
                    run(sysc.param.Tmax)
                out_host.copy_(st.height.t, non_blocking=True)
                torch.cuda.synchronize()

            wkw = workload_params(args, K)
            wkw.update(Tmax=min(K, 4), tdump=2)
            sysc_warm = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(**wkw))
            sysc, sysc_keep = sysc_warm, sysc
            job()  # untimed warm-up of the e2e path (first-use costs: module load of the operator kernels, async pool)
            sysc = sysc_keep
            dt = float("inf")
            for _ in range(3):  # best of three whole jobs (host-side jitter: allocator, Python GC, nvidia-smi teardown)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                job()
                dt = min(dt, time.perf_counter() - t0)
            plane = L * L * 8
            e2e = {"value": round(lu / dt / 1e6, 1), "unit": "MLUPS", "h2d_bytes_per_step": plane // K,
                   "d2h_bytes_per_step": (plane + 16 * 4) // K,
                   "what": f"pinned-host height -> H2D -> equilibrium! + time_loop({K} steps, mass read-back every tdump) -> D2H height; "
                           "copies amortised over the steps of the job; best of 3 jobs after one warm-up job"}
    else:
        import ctypes as C
        import torch.distributed as dist

        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            _lib.call("swalbe_dist_unique_id", raw)
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        q = sw._c_params(prm, thermal_seed=thermal_seed)
        handle = C.c_void_p()
        _lib.call("swalbe_dist_create", C.byref(handle), raw, rank, world, L, Ly_glob, C.byref(q))
        h0 = initial_height(L, rows, rank * rows, Ly_glob, workload=args.workload)
        hd = sw.Field(L, rows).set(h0)
        zero = sw.Field(L, rows)
        stream = sw._stream()
        _lib.call("swalbe_dist_set_state", handle, hd.ptr, zero.ptr, zero.ptr, None, stream)

        def set_theta():
            if moving:  # this rank's rows of the pattern; cospi.(θ) on the device
                ct = sw.cospi_field(sw.Field(L, rows).set(theta_pattern(L, rows, rank * rows, Ly_glob)))
                _lib.call("swalbe_dist_set_theta", handle, ct.ptr, stream)

        def dist_run(n, s0):
            for t0, cnt, move in (segments(s0, n) if moving else [(s0, n, False)]):
                _lib.call("swalbe_dist_time_loop", handle, cnt, t0, stream)
                if move:
                    _lib.call("swalbe_dist_shift_theta", handle, 1, 1, stream)

        set_theta()
        sampler = early_sampler
        dist_run(W, 0)
        torch.cuda.synchronize()
        dist.barrier()
        l0 = lib.swalbe_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler as clocks:
            e0.record()
            dist_run(K, W)
            e1.record()
            torch.cuda.synchronize()
        dist.barrier()
        launches = int(lib.swalbe_launch_count() - l0)
        sampler.stop()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        _lib.call("swalbe_dist_get_state", handle, hd.ptr, None, None, None, stream)
        msum = torch.tensor([hd.t.sum().item(), float(h0.sum())], device="cuda", dtype=torch.float64)
        dist.all_reduce(msum)
        mass_drift = abs(msum[0].item() - msum[1].item()) / msum[1].item()
        lu = L * Ly_glob * K
        lazy_mlups = None
        if not args.no_e2e:
            # e2e at N GPUs: per-rank pinned-host slab in, K steps with halo exchange, per-rank slab out
            h_host = torch.from_numpy(np.ascontiguousarray(h0.transpose())).pin_memory()
            out_host = torch.empty_like(h_host).pin_memory()
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            hd.t.copy_(h_host, non_blocking=True)
            _lib.call("swalbe_dist_set_state", handle, hd.ptr, zero.ptr, zero.ptr, None, stream)
            set_theta()
            dist_run(K, 0)
            _lib.call("swalbe_dist_get_state", handle, hd.ptr, None, None, None, stream)
            out_host.copy_(hd.t, non_blocking=True)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            plane = L * rows * 8
            e2e = {"value": round(lu / dt.item() / 1e6, 1), "unit": "MLUPS", "h2d_bytes_per_step": plane * world // K,
                   "d2h_bytes_per_step": plane * world // K,
                   "what": f"per-rank pinned-host slab -> H2D -> {K} fused steps with NCCL halos -> D2H slab; max over ranks"}
        _lib.call("swalbe_dist_destroy", handle)

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return
    mlups = lu / (ms * 1e-3) / 1e6
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kernel_ms = ms / K  # one fused kernel per step (plus two thin edge-strip launches per step at N > 1)
    achieved = B_ALG * L * rows / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}_{L}")
    except Exception:
        pass
    line = {
        "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args), "impl": "b200",
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
                     "frac_of_nominal_8TBs": round(achieved / 8000.0, 4), "per_gpu": True,
                     "alg_bytes_per_launch": B_ALG * L * rows,
                     # what the kernel really moves (ncu DRAM bytes of one launch / the same live kernel time): the
                     # tau = 1 kernel never reads the old populations, so it moves ~120 B/LU against the 144 B convention
                     "dram_gbs": round(traffic / (kernel_ms * 1e-3) / 1e9, 1) if traffic and world == 1 else None,
                     "dram_frac": round(traffic / (kernel_ms * 1e-3) / 1e9 / peak, 4) if traffic and world == 1 else None},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
        "mass_drift_rel": mass_drift,
    }
    if lazy_mlups is not None:
        line["moments_only_mlups"] = round(lazy_mlups, 1)  # separate row: populations written on the last step only
    if not args.no_cpu_baseline:
        from oracle import oracle_c as oc

        threads = oc.max_threads()
        m, dt = cpu_baseline_run(1024, 4, threads, workload=args.workload)
        nst = max(4, min(400, int(12.0 / (dt / 4) / 4)))
        mN, dtN = cpu_baseline_run(2048, nst, threads, workload=args.workload)
        m1, dt1 = cpu_baseline_run(1024, max(2, min(40, int(8.0 * m / threads / 1.05 + 1))), 1, workload=args.workload)
        line["cpu_baseline"] = {"value": round(mN, 2), "unit": "MLUPS", "cores": threads, "kind": "port",
                                "sample": f"2048x2048 sample of the workload, {nst} steps ({dtN:.1f} s), C restatement of the "
                                          f"reference CPU path with OpenMP x{threads}; single thread (as Julia runs it): "
                                          f"{m1:.2f} MLUPS on 1024^2"}
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
