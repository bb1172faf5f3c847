"""Small-lattice loops: us/step of the same time loop called repeatedly (1st call: plain launches, 2nd: capture +
launch, 3rd on: replay of the library's CUDA graph), marching kernel vs tile kernel, graphs on and off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import swalbe_b200 as sw

N = 1000
for L in (32, 100, 256, 512, 1024):
    for tm, gr in ((0, 0), (10**9 if L <= 512 else 0, 0), (0, 1), (10**9 if L <= 512 else 0, 1)):
        os.environ["SWALBE_TILE_MAX"], os.environ["SWALBE_GRAPH"] = str(tm), str(gr)
        sysc = sw.SysConst(Lx=L, Ly=L, param=sw.Taumucs())
        st = sw.Sys(sysc, "GPU")
        i = np.arange(L)[:, None]; j = np.arange(L)[None, :]
        st.height.set(np.asfortranarray(1 + 1e-3 * np.sin(2 * np.pi * i / L) * np.sin(2 * np.pi * j / L)))
        out = []
        for rep in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter(); sw.fused_steps(st, sysc, N, skip_aux=True); torch.cuda.synchronize()
            out.append((time.perf_counter() - t0) / N * 1e6)
        print(f"L={L} tile={'on' if tm else 'off'} graph={'on' if gr else 'off'}: us/step per call " + " ".join(f"{v:.2f}" for v in out), flush=True)
