"""tools/soak_shapes.py -- soak test of the cooperative dispersion kernels: every column of several large random models
(Rayleigh and Love, phase + group, overtones, water layer) is solved by one thread per column and by 8/16/32/128 lanes
per column; all outputs must be bit-identical.  (The thread-per-column kernel is the one the parity tests pin to the
oracle; this widens the comparison to ~10^5 columns per case.)"""
import sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi, synth
capi.init(0)
dev = torch.device('cuda', 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
bad = 0
for case, (nx, ny, nz, ncell, water, raylov, pg, nm, npd) in enumerate([
        (192, 160, 40, 900, 0.0, 1, 1, 0, 12), (160, 128, 60, 2500, 0.0, 1, 0, 2, 10), (128, 128, 30, 400, 1.3, 1, 1, 0, 8),
        (160, 160, 50, 1500, 0.0, 0, 1, 2, 10), (128, 96, 80, 3000, 0.7, 0, 0, 0, 16)]):
    grid = synth.make_grid(nx, ny, nz, waterDepth=water)
    pts, par = synth.generate_model(grid, ncell, 9000 + case)
    freqs = 1.0 / np.geomspace(0.3, 25.0, npd)
    opts = capi.disp_opts(raylov=raylov, phaseGroup=pg, nmodes=nm)
    nout = npd * max(nm, 1)
    ncell_t = nx * ny * nz
    d_vp = torch.empty(ncell_t, dtype=torch.float64, device=dev); d_vs = torch.empty_like(d_vp); d_rho = torch.empty_like(d_vp)
    d_sid = torch.empty(ncell_t, dtype=torch.int32, device=dev)
    outs = []
    for mode, lanes in ((1, 0), (2, 8), (2, 16), (2, 32), (2, 128)):
        capi.set_k2_mode(mode); capi.set_k2_lanes(lanes)
        d_pv = torch.zeros(nx * ny * nout, dtype=torch.float64, device=dev); d_gv = torch.zeros_like(d_pv)
        d_ie = torch.zeros(nx * ny, dtype=torch.int32, device=dev); d_fl = torch.zeros(2, dtype=torch.int32, device=dev)
        capi.set_nuclei_batch(*capi.pack_models([(pts, par)]))
        capi.forward_batch_dev(grid, 1, freqs, opts, d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), d_sid.data_ptr(),
                               d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), s)
        torch.cuda.synchronize()
        outs.append((d_pv, d_gv, d_ie))
    for k, o in enumerate(outs[1:]):
        same = all(torch.equal(a, b) for a, b in zip(outs[0], o))
        bad += 0 if same else 1
        print(f"case {case} ({nx*ny} columns, raylov={raylov}, group={pg}, modes={nm}, water={water}) shape {k}: {'identical' if same else 'MISMATCH'}"
              f"  ierr!=0: {int((outs[0][2] != 0).sum())}")
capi.set_k2_mode(0); capi.set_k2_lanes(0)
print("soak:", "ok" if bad == 0 else f"{bad} MISMATCHES")
sys.exit(1 if bad else 0)
