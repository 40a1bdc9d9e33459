"""tools/fm2d_bench.py -- the device eikonal solver (k6_fm2d.cuh) next to the oracle's restatement on the host cores,
for example1's geometry (101 x 101 nodes, 11 periods x 8 sources = 88 problems) and for a large batch.  One JSON line each."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mctomo_b200 import capi
import oracle_lib as orc
from concurrent.futures import ThreadPoolExecutor
from test_gpu_fm2d import _maps

capi.init(0)
ncores = os.cpu_count() or 1
for name, n, nmaps, nsrc, nrc in [("example1", 101, 11, 8, 8), ("large batch", 101, 40, 50, 50), ("256 x 256", 256, 40, 24, 24)]:
    rng = np.random.default_rng(1)
    vel = _maps(nmaps, n, n, 2)
    dx = 10.0 / (n - 1)
    src = rng.uniform(-4.5, 4.5, (nsrc, 2)); rcv = rng.uniform(-4.5, 4.5, (nrc, 2))
    srs = np.ones((nsrc, nrc), np.int32)
    o = capi.fm2d_opts()
    capi.fm2d_times(src, rcv, srs, vel, -5.0, -5.0, dx, dx, o)  # warm-up (allocations)
    capi.reset_stats()
    t = time.time()
    tt, _ = capi.fm2d_times(src, rcv, srs, vel, -5.0, -5.0, dx, dx, o)
    gpu_s = time.time() - t
    st = capi.fm2d_stats()
    # host: one period per task, as the reference's OpenMP loop over periods; a bounded sample of the periods
    sample = list(range(min(nmaps, ncores)))
    t = time.time()
    with ThreadPoolExecutor(ncores) as ex:
        res = list(ex.map(lambda m: orc.fm2d_times(src, rcv, srs, vel[m], -5.0, -5.0, dx, dx), sample))
    cpu_s = time.time() - t
    same = all(np.array_equal(res[k][1], tt[m]) for k, m in enumerate(sample))
    nprob = nmaps * nsrc
    print(json.dumps({"what": "fm2d travel times (modrays, phase-velocity data)", "case": name, "grid": [n, n], "periods": nmaps, "sources": nsrc,
                      "receivers": nrc, "problems": nprob, "gpu_s_host_to_host": gpu_s, "gpu_problems_per_s": nprob / gpu_s,
                      "nodes_accepted": st["accepted"], "gpu_nodes_per_s": st["accepted"] / gpu_s,
                      "cpu": {"cores": ncores, "threads_used": len(sample), "sample_periods": len(sample), "s": cpu_s,
                              "problems_per_s": len(sample) * nsrc / cpu_s, "kind": "port (oracle/fm2d_ref.c)"},
                      "bit_identical_to_port": bool(same)}), flush=True)
