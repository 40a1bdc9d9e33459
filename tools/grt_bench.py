"""tools/grt_bench.py -- throughput of the generalized R/T kernel (k5_grt.cuh) on a grid whose every column has a
low-velocity zone (what `program modelling` meets when the model is not monotone), next to the oracle's restatement on
the host cores (bounded sample).  Prints one JSON line."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi, synth
import oracle_lib as orc
from concurrent.futures import ThreadPoolExecutor

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
raylov = int(sys.argv[2]) if len(sys.argv) > 2 else 1
capi.init(0)
grid = synth.make_grid(nx, nx, 40)
pts, par = synth.generate_model(grid, 300, 1002)
vp, vs, rho, sid = [np.zeros(grid.shape) for _ in range(3)] + [np.zeros(grid.shape, np.int32)]
orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), vp, vs, rho, sid)
rho[:] = 1.74 * (1.73 * vs) ** 0.25; vp[:] = 1.73 * vs
rng = np.random.default_rng(9)
k0 = rng.integers(8, 20, size=(nx, nx))
for i in range(nx):
    for j in range(nx):
        k = k0[i, j]
        vs[i, j, k:k + 5] = vs[i, j, 0] * 0.85
vp[:] = 1.73 * vs; rho[:] = 1.74 * vp ** 0.25
freqs = synth.example1_freqs()
opts = capi.disp_opts(raylov=raylov, phaseGroup=0, nmodes=0)
win = (1, nx, 1, nx)
capi.set_grt(True, orc.GRT_PAR_LIKELIHOOD)
capi.reset_stats()
capi.surf_dispersion(vp, vs, rho, grid, win, freqs, opts, check=False)  # warm-up
capi.reset_stats()
t = time.time()
pv, gv, ie, inval, rc = capi.surf_dispersion(vp, vs, rho, grid, win, freqs, opts, check=False)
gpu_s = time.time() - t
st = capi.grt_stats()
# host: the oracle on a sample of the columns, all cores
ncores = os.cpu_count() or 1
sample = [(i, j) for i in range(0, nx, max(1, nx // 8)) for j in range(0, nx, max(1, nx // 8))][: 4 * ncores]
def one(ij):
    i, j = ij
    n, (th, a, b, r) = orc.convert_column(vp[i, j], vs[i, j], rho[i, j], grid.dz)
    ierr, p, g, cnt = orc.grt_modes(th, a, b, r, freqs, modetype=raylov, phaseGroup=0, par=orc.GRT_PAR_LIKELIHOOD, math_mode=orc.LIBM)
    return ierr, p, cnt, (i, j)
t = time.time()
with ThreadPoolExecutor(ncores) as ex:
    res = list(ex.map(one, sample))
cpu_s = time.time() - t
same = all(r[0] == ie[r[3]] and np.abs(np.where(r[1] < 99, r[1], 0) - np.where(pv[r[3]] < 99, pv[r[3]], 0)).max() < 1e-5 for r in res)
print(json.dumps({"what": "generalized R/T branch, every column with a low-velocity zone", "grid": [nx, nx, 40], "np": len(freqs), "raylov": raylov,
                  "columns": st["columns"], "ierr1_columns": int((ie == 1).sum()), "gpu_s_host_to_host": gpu_s,
                  "gpu_columns_per_s": st["columns"] / gpu_s, "secfun_evals": st["secfun"], "interface_steps": st["interface_steps"],
                  "secfun_evals_per_s": st["secfun"] / gpu_s,
                  "cpu": {"cores": ncores, "sample_columns": len(sample), "s": cpu_s, "columns_per_s": len(sample) / cpu_s, "kind": "port (oracle/grt_ref.c, libm)"},
                  "sample_within_1e-5": bool(same)}))
