"""tools/latency_bench.py -- latency of a proposal-sized dispersion call: thread-per-column vs warp-per-column K2 (device-resident inputs)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi, synth
capi.init(0)
dev = torch.device('cuda', 0)
grid, pts, par, freqs = synth.config(sys.argv[1] if len(sys.argv) > 1 else "C2")
opts = capi.disp_opts(raylov=1, phaseGroup=0, nmodes=0)
ncell = grid.nx * grid.ny * grid.nz
d_vp = torch.empty(ncell, dtype=torch.float64, device=dev); d_vs = torch.empty_like(d_vp); d_rho = torch.empty_like(d_vp)
d_sid = torch.empty(ncell, dtype=torch.int32, device=dev)
nout = len(freqs)
d_pv = torch.empty(grid.nx * grid.ny * nout, dtype=torch.float64, device=dev); d_gv = torch.empty_like(d_pv)
d_ie = torch.empty(grid.nx * grid.ny, dtype=torch.int32, device=dev); d_fl = torch.zeros(2, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
capi.set_nuclei_batch(*capi.pack_models([(pts, par)]))
capi.forward_batch_dev(grid, 1, freqs, opts, d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), d_sid.data_ptr(), d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), s)
torch.cuda.synchronize()
ref = d_pv.clone()
for (wx, wy) in [(4, 4), (10, 10), (20, 20), (32, 32), (64, 64), (grid.nx, grid.ny)]:
    wx = min(wx, grid.nx); wy = min(wy, grid.ny)
    win = (1, wx, 1, wy)
    out = []
    for mode in (1, 0):   # 1 = one thread per column, 0 = automatic (lane-cooperative below the GPU's capacity)
        capi.set_k2_mode(mode)
        for _ in range(2):
            capi.surf_dispersion_dev(d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), grid, win, freqs, opts, d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), s)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            capi.surf_dispersion_dev(d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), grid, win, freqs, opts, d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), s)
        b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / 3)
        capi.set_profiling(True); capi.kernel_times(reset=True)
        capi.surf_dispersion_dev(d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), grid, win, freqs, opts, d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), s)
        kt = capi.kernel_times(reset=True); capi.set_profiling(False)
        brk = f"[K2 {kt['k2_ms']:.2f} ms, other kernels {kt['other_ms']:.2f} ms]"
        res = d_pv[: wx * wy * nout].clone()
        if mode == 1: r1 = res
    same = bool(torch.equal(r1, res))
    print(f"{wx*wy:6d} columns: thread/col {out[0]:8.2f} ms   auto (G lanes/col) {out[1]:8.2f} ms   speedup {out[0]/out[1]:5.2f}x  identical={same}  auto: {brk}")
capi.set_k2_mode(0)
