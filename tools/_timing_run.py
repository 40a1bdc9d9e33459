import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi, synth
capi.init(0)
dev = torch.device('cuda', 0)
grid, pts, par, freqs = synth.config("C2")
opts = capi.disp_opts(raylov=1, phaseGroup=0, nmodes=0)
ncell = grid.nx * grid.ny * grid.nz
d_vp = torch.empty(ncell, dtype=torch.float64, device=dev); d_vs = torch.empty_like(d_vp); d_rho = torch.empty_like(d_vp)
d_sid = torch.empty(ncell, dtype=torch.int32, device=dev)
nout = len(freqs)
d_pv = torch.empty(grid.nx * grid.ny * nout, dtype=torch.float64, device=dev); d_gv = torch.empty_like(d_pv)
d_ie = torch.empty(grid.nx * grid.ny, dtype=torch.int32, device=dev); d_fl = torch.zeros(2, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
capi.set_nuclei_batch(*capi.pack_models([(pts, par)]))
capi.set_k2_mode(1)
capi.forward_batch_dev(grid, 1, freqs, opts, d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), d_sid.data_ptr(), d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), s)
torch.cuda.synchronize()
capi.set_k2_mode(2); capi.set_k2_lanes(32)
capi.surf_dispersion_dev(d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), grid, (1, 20, 1, 20), freqs, opts, d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), s)
torch.cuda.synchronize()
