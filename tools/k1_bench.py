"""tools/k1_bench.py -- K1 timing: tree walk per node vs culled brute force per column (device-resident, one model)."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from mctomo_b200 import capi, synth
capi.init(0)
dev = torch.device('cuda', 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
for name in sys.argv[1:] or ["C2", "C1", "C3"]:
    grid, pts, par, freqs = synth.config(name)
    ncell = grid.nx * grid.ny * grid.nz
    bufs = [torch.zeros(ncell, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.zeros(ncell, dtype=torch.int32, device=dev)]
    res = []
    for mode in (1, 3, 0):
        capi.set_k1_mode(mode)
        for b in bufs: b.zero_()
        for _ in range(2):
            capi.voronoi_to_grid_dev(pts, par, grid, grid.cover_box(), *[b.data_ptr() for b in bufs], s)
        torch.cuda.synchronize()
        capi.set_profiling(True); capi.kernel_times(reset=True)
        for _ in range(5):
            capi.voronoi_to_grid_dev(pts, par, grid, grid.cover_box(), *[b.data_ptr() for b in bufs], s)
        kt = capi.kernel_times(reset=True); capi.set_profiling(False)
        ms = kt["k1_ms"] / 5
        res.append([b.clone() for b in bufs])
        print(f"{name} mode {mode}: {ms:8.3f} ms  {ncell/ms/1e6:8.2f} G nodes/s  {28.0*ncell/ms/1e6:8.1f} GB/s", flush=True)
    print("   identical:", all(torch.equal(a, b) for a, b in zip(res[0], res[1])) and all(torch.equal(a, b) for a, b in zip(res[0], res[2])))
