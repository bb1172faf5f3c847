#!/usr/bin/env python
"""Throughput of the fused thin-film D2Q9 LBM step on B200 -- the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--L 8192] [--workload film|thermal]

One "step" = one pass of the hot path (src/simulate.jl:15-22) over the whole lattice.  Default workload: the
8192 x 8192 thin film the north_star target is quoted on (tau = 1, Taumucs defaults, flat film + perturbation,
SURVEY.md 8d C5/C4 geometry); at N > 1 every rank owns an L x L slab (weak scaling, global lattice L x N*L) and the
ranks exchange halo rows as peer-memory stores over NVLink (`halo_transport` in the line says which transport ran; NCCL
send/recv where the ranks cannot map each other's memory).  Prints ONE JSON line (rank 0).

  value     MLUPS with the state resident in HBM, CUDA-event timed, max over ranks.  The timed steps are the same at every
            N: one fused step kernel per step, no materialisation of the reference's intermediate fields (the last
            step of a user-level call also stores feq/pressure/h∇p/slip/F: that launch is timed separately, as
            `materialise_step_ms`)
  parity_vs_1gpu   after the timed region: a 2048 x (256 N) lattice stepped 8 times by the N-rank slab runtime and by the
            single-GPU loop on rank 0 -- film, thermal (seeded) and thermal + moving contact-angle pattern -- gathered
            and compared BITWISE (height, velocities, populations)
  e2e       MLUPS of a whole user-level job through the public API (run_host; DistSim.time_loop_host at N > 1): pinned-host
            initial height -> H2D -> time_loop (K steps, fused kernels, mass of every dump step, every field materialised
            on return) -> D2H of the final height into pinned host memory; wall clock around the job, both copies inside
            it -- they travel in row bands behind the first and ahead of the last steps (swalbe_time_loop_host)
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


class NvmlSampler:
    """SM clock, power and throttle reasons of ONE GPU, sampled every 20 ms from a thread of this process through NVML
    (nvidia_ml_py) -- on a busy 8-GPU box the nvidia-smi process of round 1 did not deliver a single line in time.
    summary() keeps the samples that fall inside the timed window.  Same interface as ClockSampler (the fallback)."""

    def __init__(self, index=0, enabled=True):
        self.rows, self.t0, self.t1, self.ok, self._stop = [], None, None, False, threading.Event()
        if not enabled:
            return
        try:
            import pynvml as nv
            import torch

            nv.nvmlInit()
            h = None
            try:  # CUDA ordinal -> NVML handle through the UUID (CUDA_VISIBLE_DEVICES may renumber the devices)
                h = nv.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(index).uuid))
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(index)
            self.nv, self.h = nv, h
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._sample()
            self.ok = True
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.ok = False

    def _sample(self):
        nv, h = self.nv, self.h
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                          nv.nvmlDeviceGetPowerUsage(h) / 1000.0, int(reasons)))

    def _pump(self):
        while not self._stop.wait(0.02):
            try:
                self._sample()
            except Exception:
                break

    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop.set()

    def summary(self):
        self.stop()
        if not self.ok or not self.rows:
            return None
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= self.t1 + 0.03]
        window = "timed region"
        if not inside:
            mid = 0.5 * (self.t0 + self.t1)
            inside, window = [min(self.rows, key=lambda r: abs(r[0] - mid))], "nearest sample"
        sm = sorted(r[1] for r in inside)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        reasons = [n for n, b in bits.items() if any(r[3] & b for r in inside)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": reasons,
                "power_w_max": max(r[2] for r in inside), "samples": len(sm), "window": window, "source": "nvml"}


def make_sampler(index, enabled):
    s = NvmlSampler(index, enabled)
    return s if (s.ok or not enabled) else ClockSampler(index, enabled)


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
                "power_w_max": max(float(r[2]) for r in inside), "samples": len(sm), "window": window, "source": "nvidia-smi"}


def host_threads():
    """Threads the CPU legs use: every core this process may run on.  Passed EXPLICITLY to the oracle's OpenMP loops --
    torch.distributed.run exports OMP_NUM_THREADS=1, which made the N > 1 reference arm of round 1 a 1-core number."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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

    threads = host_threads()
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
    rows = rows_per_rank(args)
    tau = getattr(args, "tau", 1.0)
    return {"workload": f"thin-film D2Q9 LBM step, {args.L}x{rows} per GPU, tau={tau:g}, {what}",
            "grid": [args.L, rows * args.gpus], "per_gpu_grid": [args.L, rows], "bytes_per_update_alg": B_ALG,
            "decomposition": "row slabs along y, NCCL send/recv halos" if args.gpus > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (>= 1.5 GB touched per step vs 126 MB L2)"}


def rows_per_rank(args):
    """rows of the lattice one rank owns: --rows if given, else L (weak scaling / single GPU) or L / N (strong)"""
    if getattr(args, "rows", 0):
        return args.rows
    weak = args.gpus == 1 or getattr(args, "scaling", "weak") == "weak"
    return args.L if weak else args.L // args.gpus


PARITY_L, PARITY_ROWS, PARITY_STEPS, PARITY_STEP0 = 2048, 256, 8, 94  # (steps 94..101 straddle a substrate move at 98)


def parity_vs_single_gpu(sw, _lib, rank, world):
    """N ranks == 1 GPU, bit for bit (north_star: "1/2/4/8 GPUs must give identical fields").  A 2048 x (256 N) lattice
    is stepped 8 times by the slab runtime on all ranks and by the single-GPU fused loop on rank 0, for the film, the
    thermal (seeded Philox noise) and the thermal + moving contact-angle configurations; height, velocities and the
    nine populations are gathered on rank 0 and compared with torch.equal semantics (number of differing values)."""
    import argparse as ap

    import torch

    from swalbe_b200.dist import DistSim, broadcast_unique_id_torch

    L, rows, Ly = PARITY_L, PARITY_ROWS, PARITY_ROWS * world
    out = {}
    for wl in ("film", "thermal", "thermal_moving"):
        a = ap.Namespace(workload=wl)
        prm = sw.Taumucs(**workload_params(a, PARITY_STEPS))
        sysc = sw.SysConst(Lx=L, Ly=Ly, param=prm)
        seed = 1234 if wl != "film" else None
        moving = wl == "thermal_moving"
        # (a fresh NCCL communicator per case: a DistSim owns its communicator and destroys it on close)
        sim = DistSim(sysc, rank, world, broadcast_unique_id_torch() if world > 1 else None, thermal_seed=seed)
        h = sw.Field(L, rows).set(initial_height(L, rows, rank * rows, Ly, workload=wl))
        ux, uy, f = sw.Field(L, rows), sw.Field(L, rows), sw.Field(L, rows, 9)
        sim.set_state(h, ux, uy)
        if moving:
            sim.set_theta(sw.cospi_field(sw.Field(L, rows).set(theta_pattern(L, rows, rank * rows, Ly))))
        for t0, cnt, move in (segments(PARITY_STEP0, PARITY_STEPS) if moving else [(PARITY_STEP0, PARITY_STEPS, False)]):
            sim.time_loop(cnt, t0)
            if move:
                sim.shift_theta(1, 1)
        sim.get_state(h, ux, uy, f)
        mine = torch.cat([h.t[None], ux.t[None], uy.t[None], f.t], dim=0).contiguous()  # (12, rows, L)
        if world > 1:
            import torch.distributed as dist

            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
        else:
            parts = [mine]
        sim.close()
        if rank == 0:
            got = torch.cat(parts, dim=1)  # (12, Ly, L)
            st = sw.Sys(sysc, "GPU", kind="thermal" if seed is not None else "simple")
            st.height.set(initial_height(L, Ly, 0, Ly, workload=wl))
            kw = dict(thermal_seed=seed, pressure_variant=_lib.PRESSURE_POWER_BROAD, skip_aux=True)
            if moving:
                th, inp = sw.Field(L, Ly).set(theta_pattern(L, Ly, 0, Ly)), sw.Field(L, Ly)
                inp.set(th)
                for t0, cnt, move in segments(PARITY_STEP0, PARITY_STEPS):
                    sw.fused_steps(st, sysc, cnt, step0=t0, θ=th, **kw)
                    if move:
                        sw.move_substrate(th, inp, t0 + cnt, TMOVE)
            else:
                sw.fused_steps(st, sysc, PARITY_STEPS, step0=PARITY_STEP0, **kw)
            want = torch.cat([st.height.t[None], st.velx.t[None], st.vely.t[None], st.fout.t], dim=0)
            ndiff = int((got != want).sum().item())
            moved = int((want[0] != torch.from_numpy(initial_height(L, Ly, 0, Ly, workload=wl).transpose().copy()).cuda()).sum().item())
            out[wl] = "bitwise" if (ndiff == 0 and moved > 0) else f"{ndiff} values differ" if ndiff else "fields did not change"
            del st, got, want
        del parts, mine
        torch.cuda.empty_cache()
    if rank != 0:
        return None
    verdict = "bitwise" if all(v == "bitwise" for v in out.values()) else "; ".join(f"{k}: {v}" for k, v in out.items())
    return {"verdict": verdict, "cases": out, "grid": [L, Ly], "steps": PARITY_STEPS, "ranks": world,
            "fields": "height, velx, vely, fout (9 planes)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--L", type=int, default=8192)
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: L; e.g. --L 32768 --rows 4096: C5 weak slabs)")
    ap.add_argument("--tau", type=float, default=1.0, help="relaxation time (every BASELINE config has tau = 1; tau != 1 runs "
                                                            "the general kernels that read the old populations)")
    ap.add_argument("--workload", default="film", choices=["film", "thermal", "thermal_moving", "droplet", "spinodal"],
                    help="film: C5 flat film + sine (default); thermal: film + Philox noise; thermal_moving: the complete C4 "
                         "(noise + contact-angle pattern moved by (1,1) every 98 steps, n=3,m=2,hmin=0.07,mu=1/12); droplet: C2 spherical cap, "
                         "n=3,m=2,hmin=0.07 (use --L 1024); spinodal: C3 random film, n=3,m=2,hmin=0.07,gamma=0.01 (--L 4096)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = L x L per rank (default, the contract), strong = L x L in total (L/N rows per rank)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--check-e2e", action="store_true", help="after each e2e job, redo it as copy + equilibrium! + time_loop + "
                                                             "copy and compare the heights bitwise (untimed runs only)")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank-vs-1-GPU bitwise leg")
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
    if rank == 0:
        ge.build()
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import swalbe_b200 as sw
    from swalbe_b200 import _lib

    lib = _lib.load()
    sampler = make_sampler(local_rank, enabled=rank == 0)  # clocks of rank 0's GPU, sampled during the timed region
    L, K, W = args.L, args.steps, args.warmup
    wkw = workload_params(args, K)
    if args.tau != 1.0:
        wkw.update(τ=args.tau)
    prm = sw.Taumucs(**wkw)
    rows = rows_per_rank(args)   # rows per rank
    Ly_glob = rows * world
    sysc = sw.SysConst(Lx=L, Ly=rows, param=prm)
    thermal_seed = 1234 if args.workload in ("thermal", "thermal_moving") else None
    moving = args.workload == "thermal_moving"
    e2e = None
    extra = {}

    if world == 1:
        st = sw.Sys(sysc, "GPU", kind="thermal" if thermal_seed is not None else "simple")
        h0 = initial_height(L, rows, workload=args.workload)
        st.height.set(h0)
        th = inp = None
        if moving:
            th, inp = sw.Field(L, rows).set(theta_pattern(L, rows)), sw.Field(L, rows)
            inp.set(th)
        state = {"started": False}

        def run(n, s0=0, aux=False, host_in=None, host_out=None, **kw):
            """n steps; `aux`: the last launch also materialises the reference's intermediate fields (user-level call);
            host_in / host_out: pinned host planes the first segment starts from / the last one leaves the height in"""
            segs = segments(s0, n) if moving else [(s0, n, False)]
            for q, (t0, cnt, move) in enumerate(segs):
                sw.fused_steps(st, sysc, cnt, thermal_seed=thermal_seed, step0=t0, θ=th, skip_aux=not (aux and q == len(segs) - 1),
                               pressure_variant=_lib.PRESSURE_POWER_BROAD, moments_consistent=state["started"],
                               host_in=host_in if q == 0 else None, host_out=host_out if q == len(segs) - 1 else None, **kw)
                state["started"] = True
                if move:
                    sw.move_substrate(th, inp, t0 + cnt, TMOVE)
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
        lu = L * rows * K
        # the launch that ends a user-level call: the same step + feq/vsq/pressure/h∇p/slip/F and the second population
        # copy written out (38 planes instead of 12) -- timed on its own, never part of `value`
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        run(1, W + K, aux=True)
        e4.record()
        for q in range(3):
            run(1, W + K + 1 + q, aux=True)
        e5.record()
        torch.cuda.synchronize()
        extra["materialise_step_ms"] = round(e4.elapsed_time(e5) / 3, 4)
        lazy_mlups = None
        if args.tau == 1.0:
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

            def job(sc):
                st.velx.t.zero_(); st.vely.t.zero_()
                if thermal_seed is None:
                    sw.run_host(sc, h_host, out_host, state=st)  # height from the host, time_loop, height to the host
                else:
                    state["started"] = False
                    run(sc.param.Tmax, aux=True, host_in=h_host, host_out=out_host)
                torch.cuda.synchronize()
                if args.check_e2e:  # the streamed job against the plain sequence copy, equilibrium!, time_loop, copy
                    got = out_host.clone()
                    st.height.t.copy_(h_host, non_blocking=True)
                    st.velx.t.zero_(); st.vely.t.zero_()
                    sw.equilibrium(st, sc)
                    if thermal_seed is None:
                        sw.time_loop(sc, st)
                    else:
                        state["started"] = False
                        run(sc.param.Tmax, aux=True)
                    extra["e2e_equals_plain_sequence"] = bool(torch.equal(got, st.height.t.cpu()))

            wk2 = dict(wkw)
            wk2.update(Tmax=min(K, 4), tdump=2)
            job(sw.SysConst(Lx=L, Ly=rows, param=sw.Taumucs(**wk2)))  # untimed warm-up of the e2e path (first-use costs)
            dt = float("inf")
            for _ in range(3):  # best of three whole jobs (host-side jitter: allocator, Python GC)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                job(sysc)
                dt = min(dt, time.perf_counter() - t0)
            plane = L * rows * 8
            e2e = {"value": round(lu / dt / 1e6, 1), "unit": "MLUPS", "h2d_bytes_per_step": plane // K,
                   "d2h_bytes_per_step": (plane + 16 * 4) // K,
                   "what": f"run_host: pinned-host height -> H2D -> time_loop({K} steps, mass read-back every tdump, every field of "
                           "the state materialised on return) -> D2H height into pinned host memory, both copies inside the timed "
                           "region, travelling in row bands behind / ahead of the first / last steps (swalbe_time_loop_host); "
                           "best of 3 jobs after one warm-up job"}
        del st
        torch.cuda.empty_cache()
    else:
        import torch.distributed as dist

        from swalbe_b200.dist import DistSim, broadcast_unique_id_torch

        sysg = sw.SysConst(Lx=L, Ly=Ly_glob, param=prm)
        sim = DistSim(sysg, rank, world, broadcast_unique_id_torch(), thermal_seed=thermal_seed)
        h0 = initial_height(L, rows, rank * rows, Ly_glob, workload=args.workload)
        hd = sw.Field(L, rows).set(h0)
        zero = sw.Field(L, rows)
        sim.set_state(hd, zero, zero)

        th_host = None
        if moving:  # this rank's rows of the contact-angle pattern, built once on the host (pinned): an INPUT of the job
            th_host = torch.from_numpy(np.ascontiguousarray(theta_pattern(L, rows, rank * rows, Ly_glob).transpose())).pin_memory()
        th_dev = sw.Field(L, rows) if moving else None

        def set_theta():
            if moving:  # H2D of the pattern, cospi.(θ) on the device, ghost rows through the halo exchange
                th_dev.t.copy_(th_host, non_blocking=True)
                sim.set_theta(sw.cospi_field(th_dev.touch()))

        def dist_run(n, s0):
            for t0, cnt, move in (segments(s0, n) if moving else [(s0, n, False)]):
                sim.time_loop(cnt, t0)
                if move:
                    sim.shift_theta(1, 1)

        set_theta()
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
        extra["halo_transport"] = "peer-memory stores over NVLink (k_halo_push)" if sim.uses_peer_memory() else "NCCL send/recv"
        if not moving:  # device time of the K-step loop on rank 0's own streams (edge strips + halo exchange + interior):
            # next to the single-GPU kernel time it shows whether the exchange is hidden behind the interior update
            extra["dist_loop_ms_per_step_rank0"] = round(sim.last_loop_ms() / K, 4)
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        sim.get_state(hd)
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
            set_theta()
            segs = segments(0, K) if moving else [(0, K, False)]
            for q, (s0_, cnt, move) in enumerate(segs):  # the slab's rows travel in bands behind / ahead of the steps
                sim.time_loop_host(cnt, host_in=h_host if q == 0 else None, host_out=out_host if q == len(segs) - 1 else None,
                                   step0=s0_)
                if move:
                    sim.shift_theta(1, 1)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            plane = L * rows * 8
            e2e = {"value": round(lu / dt.item() / 1e6, 1), "unit": "MLUPS", "h2d_bytes_per_step": plane * world * (2 if moving else 1) // K,
                   "d2h_bytes_per_step": plane * world // K,
                   "what": f"per-rank pinned-host slab -> H2D -> {K} fused steps with halo exchange -> D2H slab into pinned host "
                           "memory, both copies inside the timed region, travelling in row bands behind / ahead of the first / last "
                           "steps (swalbe_dist_time_loop_host); wall clock, max over ranks"}
        sim.close()
        del hd, zero
        torch.cuda.empty_cache()

    parity = None if args.no_parity else parity_vs_single_gpu(sw, _lib, rank, world)

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
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(f"{args.workload}_{L}" + (f"_tau{args.tau:g}" if args.tau != 1.0 else "")) if rows == L else None
    except Exception:
        pass
    clk = clocks.summary() or {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable"]}
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
        "e2e": e2e, "gpu_launches": launches, "clocks": clk,
        "mass_drift_rel": mass_drift,
    }
    line.update(extra)
    if parity is not None:
        line["parity_vs_1gpu"] = parity["verdict"]
        line["parity_detail"] = parity
    if lazy_mlups is not None:
        line["moments_only_mlups"] = round(lazy_mlups, 1)  # separate row: populations written on the last step only
    if not args.no_cpu_baseline and world == 1:  # (rank 0 at N = 1 only: the tier's contract)
        threads = host_threads()
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
