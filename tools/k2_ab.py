"""tools/k2_ab.py -- A/B of dispersion-kernel builds on the bench-default workload (C2 x 32): K2 device time with the
automatic shape and with one thread per column.  MCT_LIB selects the library build (mctomo_b200/capi.py)."""
import os
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from mctomo_b200 import capi, synth
capi.init(0)
dev = torch.device('cuda', 0)
tstream = torch.cuda.Stream(); torch.cuda.set_stream(tstream); s = tstream.cuda_stream
grid = synth.make_grid(64, 64, 40); freqs = synth.freqs(20)
batch = 32
models = [synth.generate_model(grid, 300, 1002 + b) for b in range(batch)]
pts, par, off = capi.pack_models(models)
opts = capi.disp_opts()
n = grid.nx * grid.ny * grid.nz * batch
vp, vs, rho = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
sid = torch.empty(n, dtype=torch.int32, device=dev)
ncol = grid.nx * grid.ny * batch
pv = torch.empty(ncol * 20, dtype=torch.float64, device=dev); gv = torch.empty_like(pv)
ie = torch.empty(ncol, dtype=torch.int32, device=dev); fl = torch.zeros(2 * batch, dtype=torch.int32, device=dev)
capi.set_nuclei_batch(pts, par, off)
for mode in (0, 1):
    capi.set_k2_mode(mode, -1)
    for _ in range(2):
        capi.forward_batch_dev(grid, batch, freqs, opts, vp.data_ptr(), vs.data_ptr(), rho.data_ptr(), sid.data_ptr(), pv.data_ptr(), gv.data_ptr(), ie.data_ptr(), fl.data_ptr(), s)
    torch.cuda.synchronize()
    capi.set_profiling(True); capi.kernel_times(reset=True)
    for _ in range(4):
        capi.forward_batch_dev(grid, batch, freqs, opts, vp.data_ptr(), vs.data_ptr(), rho.data_ptr(), sid.data_ptr(), pv.data_ptr(), gv.data_ptr(), ie.data_ptr(), fl.data_ptr(), s)
    kt = capi.kernel_times(reset=True); capi.set_profiling(False)
    print(f"{os.environ.get('MCT_LIB', 'default'):28s} mode {mode}: K2 {kt['k2_ms'] / 4:7.2f} ms  kernel {capi.last_launch()['kernel']}  checksum {float(pv.sum()):.6f}", flush=True)
