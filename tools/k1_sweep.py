"""tools/k1_sweep.py -- K1 on the bench workload (C2 x 32 models in one batched launch) and on C5, for a list of tile
shapes (MCT_K1_TILE = ttx,tty,seglen; read by the library at every launch).  Device time from the library's own events."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from mctomo_b200 import capi, synth
capi.init(0)
dev = torch.device('cuda', 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
which = sys.argv[1] if len(sys.argv) > 1 else "C2x32"
shapes = sys.argv[2:] or ["auto"]
if which == "C2x32":
    grid, _, _, freqs = synth.config("C2")
    models = [synth.generate_model(grid, 300, 1002 + b) for b in range(32)]
    nb = 32
else:
    grid, pts, par, freqs = synth.config(which)
    models = [(pts, par)]
    nb = 1
pts, par, off = capi.pack_models(models)
ncell = grid.nx * grid.ny * grid.nz
bufs = [torch.zeros(ncell * nb, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.zeros(ncell * nb, dtype=torch.int32, device=dev)]
capi.set_nuclei_batch(pts, par, off)
opts = capi.disp_opts(raylov=1, phaseGroup=0, nmodes=0)
f1 = freqs[:1]  # one period: the dispersion kernel is along for the ride, K1 is timed by the library's own events
d_pv = torch.zeros(grid.nx * grid.ny * nb, dtype=torch.float64, device=dev); d_gv = torch.zeros_like(d_pv)
d_ie = torch.zeros(grid.nx * grid.ny * nb, dtype=torch.int32, device=dev); d_fl = torch.zeros(2 * nb, dtype=torch.int32, device=dev)
ref = None
for mode, shape in [(3, "auto")] + [(0, sh) for sh in shapes]:
    capi.set_k1_mode(mode)
    for k in ("MCT_K1_TILE", "MCT_K1_NPT", "MCT_K1_BPS"): os.environ.pop(k, None)
    for part in shape.split("/"):  # e.g. 4,8,4/npt2/bps24
        if part.startswith("npt"): os.environ["MCT_K1_NPT"] = part[3:]
        elif part.startswith("bps"): os.environ["MCT_K1_BPS"] = part[3:]
        elif part != "auto": os.environ["MCT_K1_TILE"] = part
    for b in bufs: b.zero_()
    def run():
        capi.forward_batch_dev(grid, nb, f1, opts, *[b.data_ptr() for b in bufs], d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), s)
    for _ in range(3): run()
    torch.cuda.synchronize()
    capi.set_profiling(True); capi.kernel_times(reset=True)
    for _ in range(10): run()
    kt = capi.kernel_times(reset=True); capi.set_profiling(False)
    ms = kt["k1_ms"] / 10
    out = [b.clone() for b in bufs]
    same = True if ref is None else all(torch.equal(a, b) for a, b in zip(ref, out))
    if ref is None: ref = out
    print(f"{which} mode {mode} tile {shape:>18}: {ms*1e3:9.1f} us  {28.0*ncell*nb/ms/1e6:8.1f} GB/s  identical to mode 3: {same}", flush=True)
