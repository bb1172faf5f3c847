"""us/step of small lattices through the public fused loop, per call (first call, capture, replays), for the persistent
cluster kernel / tile kernel / marching kernel (SWALBE_CLUSTER, SWALBE_CLUSTER_SIZE, SWALBE_TILE_MAX), with and without
per-step logs (the README Rayleigh-Taylor example logs max - min every step).
   python tools/small_probe.py L [steps_per_call] [logs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swalbe_b200 as sw

L = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 100; logs = len(sys.argv) > 3 and sys.argv[3] == "logs"
sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs(g=-0.001, γ=0.0005))
st = sw.Sys(sysc, "GPU")
i = np.arange(1, L + 1)[:, None]; j = np.arange(1, L + 1)[None, :]
st.height.set(1.0 + 0.01 * np.sin(2 * np.pi * 15 * i / (L - 1)) * np.sin(2 * np.pi * 18 * j / (L - 1)))
out = []
for c in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    sw.fused_steps(st, sysc, n, skip_aux=True, log_minmax=logs)
    e1.record(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    out.append(f"{e0.elapsed_time(e1) * 1e3 / n:.2f}/{wall * 1e6 / n:.2f}")
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SWALBE_"))
print(f"L={L} steps/call={n} logs={logs} {env}: us/step device/wall per call: " + " ".join(out), flush=True)
