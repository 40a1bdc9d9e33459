"""tools/slab_k2_modes.py -- one C5 x-slab (1024 x 128 columns of the 1024 x 1024 x 80 grid, 60 periods) on one GPU: which
dispersion-kernel shape is best at the slab's size (the 8-GPU column-sharded run solves eight of these)."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi, synth
capi.init(0)
dev = torch.device('cuda', 0)
grid, pts, par, freqs = synth.config("C5")
opts = capi.disp_opts(raylov=1, phaseGroup=0, nmodes=0)
ncell = grid.nx * grid.ny * grid.nz
d_vp = torch.empty(ncell, dtype=torch.float64, device=dev); d_vs = torch.empty_like(d_vp); d_rho = torch.empty_like(d_vp)
d_sid = torch.empty(ncell, dtype=torch.int32, device=dev)
nout = len(freqs)
d_pv = torch.empty(grid.nx * grid.ny * nout, dtype=torch.float64, device=dev); d_gv = torch.empty(8, dtype=torch.float64, device=dev)
d_ie = torch.empty(grid.nx * grid.ny, dtype=torch.int32, device=dev); d_fl = torch.zeros(2, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
capi.set_nuclei_batch(*capi.pack_models([(pts, par)]))
slab = (385, 512)
ref = None
for name, mode, lanes in [("auto", 0, 0), ("thread per column", 1, 0), ("2 lanes", 2, 2), ("4 lanes", 2, 4)]:
    capi.set_k2_mode(mode); capi.set_k2_lanes(lanes)
    def run():
        capi.forward_batch_dev(grid, 1, freqs, opts, d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), d_sid.data_ptr(), d_pv.data_ptr(), d_gv.data_ptr(),
                               d_ie.data_ptr(), d_fl.data_ptr(), s, slab=slab)
    run(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); run(); run(); b.record(); torch.cuda.synchronize()
    out = d_pv.clone()
    same = True if ref is None else bool(torch.equal(out, ref))
    if ref is None: ref = out
    print(f"{name:20s} {a.elapsed_time(b)/2:8.1f} ms per slab evaluation   kernel {capi.last_launch()['kernel']}  identical {same}", flush=True)
capi.set_k2_mode(0); capi.set_k2_lanes(0)
