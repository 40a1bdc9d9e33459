"""tools/lanes_sweep.py -- dispersion latency of small column batches as a function of lanes per column
(2..32: lane groups inside a warp; 64/128: several warps per column).  Device-resident inputs; used to set
the automatic choice in launch_k2 (mct_api.cu)."""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi, synth
capi.init(0)
dev = torch.device('cuda', 0)
if len(sys.argv) > 1 and sys.argv[1] == "X":   # 256 x 256 x 40 grid, 20 periods: the hand-over to one thread per column
    grid = synth.make_grid(256, 256, 40)
    pts, par = synth.generate_model(grid, 1200, 77)
    freqs = synth.freqs(20)
else:
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


def run(win, n=3):
    args = (d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), grid, win, freqs, opts, d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), s)
    for _ in range(2):
        capi.surf_dispersion_dev(*args)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        capi.surf_dispersion_dev(*args)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


capi.set_k2_mode(2)
BIG = len(sys.argv) > 2 and sys.argv[2] == "big"
HUGE = len(sys.argv) > 2 and sys.argv[2] == "huge"
WINS = [(90, 90), (128, 128), (181, 181), (222, 222), (256, 256)] if HUGE else [(45, 45), (64, 64), (80, 80), (grid.nx, grid.ny)] if BIG else [(4, 4), (8, 8), (12, 12), (16, 16), (20, 20), (24, 24), (32, 32), (45, 45)]
LANES = (1, 2, 4, 8, 16, 0) if HUGE else (2, 4, 8, 16, 32, 0) if BIG else (32, 64, 128, 256, 0)   # 1 = one thread per column
for (wx, wy) in WINS:
    wx = min(wx, grid.nx); wy = min(wy, grid.ny)
    win = (1, wx, 1, wy)
    row = []
    base = None
    for lanes in LANES:
        capi.set_k2_mode(1 if lanes == 1 else 2)
        capi.set_k2_lanes(0 if lanes == 1 else lanes)
        ms = run(win)
        res = d_pv[: wx * wy * nout].clone()
        if base is None:
            base = res
        row.append(f"{'auto' if lanes == 0 else lanes}: {ms:6.2f}{'' if torch.equal(res, base) else ' MISMATCH'}")
    print(f"{wx*wy:5d} columns   " + "   ".join(row))
capi.set_k2_mode(0); capi.set_k2_lanes(0)
