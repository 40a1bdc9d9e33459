#!/usr/bin/env python
"""bench.py -- the hot path of MCTomo's surface-wave forward model on N B200s, next to the reference's CPU path.

One "step" = one pass of the hot path over one batch of synthetic models:
    Voronoi->grid (K1) -> vs2vp/vp2rho -> check_model -> column->layers -> modal dispersion (K2)
for `--batch` independent models (chains) of the workload grid.  Default workload = BASELINE config C2
(64x64x40 grid, 20 periods, Rayleigh fundamental phase velocity, 300 nuclei, batch of 32 forward evaluations).

  value : (column x period) dispersion solves/s, whole job, nuclei trees + all model arrays resident in HBM
          (timed with CUDA events on the launching stream, max over ranks)
  e2e   : same metric through the host-pointer C-ABI call mct_forward_eval_batch: nuclei come from host memory,
          trees are built, everything runs, the dispersion maps are copied back to pinned host memory.
  N > 1 : one process per GPU (torchrun), every rank evaluates its own batch of chains (config C4's
          "chains sharded over GPUs, no inter-GPU traffic"): weak scaling, no data-path collective.
          `--config C5` instead shards the columns of ONE model over the ranks in x-slabs and all-gathers
          the dispersion map with NCCL (config 5).

--impl reference times the reference's CPU path (the C restatement under oracle/, libm math, OpenMP over x
like the reference) on this box's host cores, for the same metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_LAYER_R, F_CALL_R = 184, 31  # nominal FP64 ops per layer step / per call, Rayleigh
F_LAYER_L, F_CALL_L = 27, 8    # Love


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--batch", type=int, default=0, help="models per GPU per step (default: 32 for C2, 8 for C4, else 1)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the column-sharded C5 block (one chain's columns over the ranks)")
    ap.add_argument("--c5-steps", type=int, default=2)
    ap.add_argument("--no-widened", action="store_true", help="skip the block on the widened path (generalized R/T columns, fm2d travel times)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        if not any(len(r) >= 9 for r in self.rows):  # region shorter than the sampling period: one immediate sample
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout
                self.rows += [[c.strip() for c in line.split(",")] for line in out.splitlines()]
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def k2_source_sha():
    """Fingerprint of everything the dispersion kernel is compiled from: ties profiles/k2_traffic.json to a build."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "mctomo_b200", "csrc")
    for f in ("k2_dispersion.cuh", "k2_rayleigh_fast.cuh", "k2_love_fast.cuh", "k2_coop.cuh", "k2_layerpar.cuh", "mct_math.h", "Makefile"):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def hbm_peak():
    """Measured HBM copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the recipe's fallback."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6500.0


def workload(args, rank, world):
    from mctomo_b200 import synth
    c = dict(synth.CONFIGS[args.config])
    batch = args.batch or {"C2": 32, "C4": 8}.get(args.config, 1)
    grid = synth.make_grid(c["nx"], c["ny"], c["nz"])
    cid = int(args.config[1:])
    rng_nc = np.random.default_rng(77 + rank)
    models = []
    for b in range(batch):
        chain = rank * batch + b
        ncells = c["ncells"] if args.config != "C4" else int(rng_nc.integers(25, 301))
        models.append(synth.generate_model(grid, ncells, 1000 + cid + chain))
    freqs = synth.example1_freqs() if args.config == "C1" else synth.freqs(c["np"])
    spec = dict(raylov=1, phaseGroup=0, nmodes=0)
    if args.config == "C3":
        spec = dict(raylov=1, phaseGroup=1, nmodes=2)  # Rayleigh pass; the Love pass of C3 is `love_spec` below
    return grid, models, freqs, batch, spec


def workload_config(args, grid, freqs, spec, ncells, batch):
    """`config` of the JSON line: the SAME dict for both arms (the driver compares them); how each arm runs the
    workload is described under that arm's own keys (`execution` here, cpu_baseline.sample for the reference)."""
    return {"workload": f"{args.config}: {grid.nx}x{grid.ny}x{grid.nz} grid, {len(freqs)} periods, Rayleigh "
                        f"{'phase+group' if spec['phaseGroup'] else 'phase'}, modes={max(spec['nmodes'], 1)}, {ncells} nuclei",
            "batch_per_gpu": batch,
            "l2": "GPU arm: a 256 MiB buffer is written between timed steps (L2 flush), and a step's model arrays exceed the "
                  "126 MB L2 at the default batch; CPU arm: not applicable"}


def cpu_sample(grid, models, freqs, spec, budget_s, nthreads):
    """Times the oracle (libm math = what the Fortran binary calls) on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    # the reference's own surfdisp96.f (oracle/_ref: translated to C by oracle/f77toc.py, gcc -O2) when it travelled
    # with the repository; else the port (libm math, what the Fortran binary calls)
    mode = orc.REFERENCE if orc.use_reference_solver() else orc.LIBM
    cpu_sample.kind = "reference" if mode == orc.REFERENCE else "port"
    solves = 0
    t_used = 0.0
    n_evals = 0
    # a bounded x-slab of each model keeps the sample within the budget on big grids
    per_eval_cols = grid.nx * grid.ny
    probe_cols = min(per_eval_cols, 64 * nthreads)
    wx = max(1, min(grid.nx, probe_cols // grid.ny))
    detail = None
    # successive x-slabs of the models, in turn, until the budget is used (slab k of every model before slab k + 1)
    work = [(k, m) for k in range(max(1, grid.nx // wx)) for m in models]
    for k, (pts, par) in work:
        x0 = k * wx
        t0 = time.perf_counter()
        vp = np.zeros(grid.shape); vs = np.zeros(grid.shape); rho = np.zeros(grid.shape); sid = np.zeros(grid.shape, np.int32)
        orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), vp, vs, rho, sid)
        t1 = time.perf_counter()
        vp, rho = orc.vs2vp_rho(vs, orc.LIBM)
        inval = orc.check_model(vs, grid)
        t2 = time.perf_counter()
        specs = spec if isinstance(spec, list) else [spec]
        for sp in specs:
            pv, gv, ie, cnt, _ = orc.surf_dispersion(vp, vs, rho, grid, (x0 + 1, x0 + wx, 1, grid.ny), freqs, math_mode=mode,
                                                     nthreads=nthreads, **sp)
            solves += wx * grid.ny * len(freqs) * max(sp["nmodes"], 1)
        t3 = time.perf_counter()
        frac = wx / grid.nx  # K1/property maps covered the whole grid: charge them pro rata to the solved slab
        t_used += (t1 - t0) * frac + (t2 - t1) * frac + (t3 - t2)
        n_evals += frac
        detail = {"kdtree_to_grid_s_full_grid": t1 - t0, "maps_check_s_full_grid": t2 - t1, "dispersion_s_slab": t3 - t2,
                  "slab_columns": wx * grid.ny}
        if t_used > budget_s:
            break
    what = ("dispersion by the reference's own surfdisp96.f (oracle/_ref: mechanical C translation, gcc -O2 -- the reference builds "
            "it without optimisation flags), kd-tree/layering/loops by the C port, OpenMP over x like the reference"
            if mode == orc.REFERENCE else "C port of the path (oracle/, libm math), OpenMP over x like the reference")
    return solves / t_used, n_evals / t_used, f"{n_evals:.2f} forward evals ({solves} solves) of the workload, {t_used:.1f} s; {what}", detail


cpu_sample.kind = "port"



def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    grid, models, freqs, batch, spec = workload(args, 0, 1)
    cores = os.cpu_count() or 1
    times = []
    rate = evr = 0.0
    sample = ""
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    wall = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cspec = [spec, dict(raylov=0, phaseGroup=1, nmodes=2)] if args.config == "C3" else spec
        r, e, sample, _ = cpu_sample(grid, models[s % len(models):] + models[:s % len(models)], freqs, cspec, per_step_budget, cores)
        if s >= args.warmup:
            times.append((r, e))
            wall.append(time.perf_counter() - t0)
    per_step_ms = statistics.mean(wall) * 1e3
    rate = statistics.mean(t[0] for t in times)
    evr = statistics.mean(t[1] for t in times)
    out = {"impl": "reference", "metric": "dispersion_solves_per_sec", "value": rate, "unit": "column*period solves/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "forward_evals_per_sec": evr,
           "config": workload_config(args, grid, freqs, spec, len(models[0][0]), batch),
           "execution": {"per_step": "bounded sample of the workload on the host CPU, all host threads"},
           "cpu_baseline": {"value": rate, "unit": "column*period solves/s", "cores": cores, "kind": cpu_sample.kind, "sample": sample},
           "e2e": {"value": rate, "unit": "column*period solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    GUARD.emit(json.dumps(out))


def c5_block(args, torch, dist, capi, dev, rank, world, stream):
    """BASELINE config 5 next to the headline: ONE chain on the 1024x1024x80 grid, 60 periods, 5000 nuclei, its columns
    sharded over the ranks in x-slabs; every rank runs mct_forward_sharded_dev (K1 + maps + check + dispersion of its slab,
    then the in-place ncclAllGather of the (60,1024,1024) map + ierr and the flag all-reduce, all inside the library).
    Strong scaling: per-rank K1/K2 device times, collective time, and the same evaluation unsharded on rank 0 for the
    efficiency.  Communicator bootstrap: the 128-byte id travels by torch.distributed.broadcast (MPI_Bcast in MCTomo)."""
    from mctomo_b200 import synth
    c = synth.CONFIGS["C5"]
    grid = synth.make_grid(c["nx"], c["ny"], c["nz"])
    freqs = synth.freqs(c["np"])
    nout = len(freqs)
    opts = capi.disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    pts, par = synth.generate_model(grid, c["ncells"], 1005)   # the same model on every rank
    off = np.array([0, len(pts)], np.int64)
    if world > 1 and not capi.comm_info()["active"]:
        capi.comm_init_torch(dist, dev)
    lo, hi, per = capi.slab_bounds(grid.nx, world, rank)
    n = grid.nx * grid.ny * grid.nz
    cols = per * world * grid.ny
    d_vp = torch.empty(n, dtype=torch.float64, device=dev); d_vs = torch.empty_like(d_vp); d_rho = torch.empty_like(d_vp)
    d_sid = torch.empty(n, dtype=torch.int32, device=dev)
    d_pv = torch.empty(cols * nout, dtype=torch.float64, device=dev); d_gv = torch.empty_like(d_pv)
    d_ie = torch.empty(cols, dtype=torch.int32, device=dev); d_fl = torch.zeros(2, dtype=torch.int32, device=dev)
    capi.set_nuclei_batch(pts, par, off)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        capi.forward_sharded_dev(grid, freqs, opts, d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), d_sid.data_ptr(),
                                 d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), stream)

    def measure(mode):
        capi.comm_set_mode(mode)
        step()  # warm-up (buffers grow, NCCL connects)
        sync()
        capi.set_profiling(True)
        capi.kernel_times(reset=True)
        capi.reset_stats()
        ms, comm_ms = [], []
        for _ in range(args.c5_steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sync()
            a.record()
            step()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
            comm_ms.append(capi.comm_last_ms())
        kt = capi.kernel_times(reset=True)
        st = capi.stats()
        capi.set_profiling(False)
        mine = torch.tensor([sum(ms) / len(ms), kt["k1_ms"] / len(ms), kt["k2_ms"] / len(ms), kt["other_ms"] / len(ms),
                             sum(comm_ms) / len(ms), float(st["n_columns_solved"]) / len(ms), float((hi - lo + 1) * grid.ny)],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, mine)
        else:
            allr = [mine]
        return allr, d_pv.clone()

    balanced = None
    if world > 1:  # the balanced division first (its maps are compared with the unsharded ones below)
        allr_b, pv_b = measure(1)
        balanced = torch.stack(allr_b).cpu().numpy()
    allr, _ = measure(0)
    allr = torch.stack(allr).cpu().numpy()
    ms_eval = float(allr[:, 0].max())
    # the same evaluation unsharded on ONE GPU (rank 0; the others wait): the strong-scaling denominator
    n1_ms = ms_eval
    if world > 1:
        if rank == 0:
            d_p1 = torch.empty(grid.nx * grid.ny * nout, dtype=torch.float64, device=dev); d_g1 = torch.empty_like(d_p1)
            d_i1 = torch.empty(grid.nx * grid.ny, dtype=torch.int32, device=dev)
            t = []
            for it in range(2):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                capi.forward_batch_dev(grid, 1, freqs, opts, d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), d_sid.data_ptr(),
                                       d_p1.data_ptr(), d_g1.data_ptr(), d_i1.data_ptr(), d_fl.data_ptr(), stream)
                b.record()
                torch.cuda.synchronize()
                t.append(a.elapsed_time(b))
            n1_ms = t[-1]
            same = bool(torch.equal(d_p1, d_pv[: d_p1.numel()]))  # gathered map == unsharded map, bit for bit
            same_b = bool(torch.equal(d_p1, pv_b[: d_p1.numel()]))
        else:
            same = same_b = True
        sync()
        t1 = torch.tensor([n1_ms], dtype=torch.float64, device=dev)
        dist.broadcast(t1, 0)
        n1_ms = float(t1.item())
    else:
        same = same_b = True
    capi.kernel_times(reset=True)
    capi.set_profiling(False)
    solves = grid.nx * grid.ny * nout
    bal = None
    if balanced is not None:
        bms = float(balanced[:, 0].max())
        bal = {"what": "mct_comm_set_mode(1): every rank grids, layers and de-duplicates the whole model, solves every n-th entry of "
                       "the sorted list of distinct columns; compact results all-gathered in place",
               "ms_per_eval": bms, "solves_per_sec": solves / (bms * 1e-3), "per_rank_k1_ms": [float(v) for v in balanced[:, 1]],
               "per_rank_k2_ms": [float(v) for v in balanced[:, 2]], "per_rank_other_kernels_ms": [float(v) for v in balanced[:, 3]],
               "per_rank_allgather_ms": [float(v) for v in balanced[:, 4]],
               "per_rank_distinct_columns_solved": [float(v) for v in balanced[:, 5]],
               "efficiency_vs_n1": n1_ms / (world * bms), "gathered_equals_unsharded": same_b}
    return {"workload": f"C5: {grid.nx}x{grid.ny}x{grid.nz} grid, {nout} periods, Rayleigh phase, {len(pts)} nuclei, ONE chain, "
                        f"x-slabs of {per} columns per rank", "scaling": "strong", "n_gpus": world, "steps": args.c5_steps,
            "ms_per_eval": ms_eval, "solves_per_sec": solves / (ms_eval * 1e-3),
            "per_rank_ms": [float(v) for v in allr[:, 0]], "per_rank_k1_ms": [float(v) for v in allr[:, 1]],
            "per_rank_k2_ms": [float(v) for v in allr[:, 2]], "per_rank_other_kernels_ms": [float(v) for v in allr[:, 3]],
            "allgather_ms": float(allr[:, 4].max()), "per_rank_allgather_ms": [float(v) for v in allr[:, 4]],
            "allgather_bytes_per_rank": int(per * grid.ny * (nout * 8 + 4)),
            "per_rank_distinct_columns_solved": [float(v) for v in allr[:, 5]], "per_rank_columns": [float(v) for v in allr[:, 6]],
            "n1_ms_per_eval": n1_ms, "efficiency_vs_n1": n1_ms / (world * ms_eval),
            "gathered_equals_unsharded": same, "balanced": bal,
            "collective": "in-place ncclAllGather (pvel f64 + ierr i32) + ncclAllReduce(MAX) of 2 flags, inside libmctomo_b200.so "
                          f"(NCCL {capi.comm_info()['nccl_version']})" if world > 1 else "none (one rank)"}


class StdoutGuard:
    """Keeps stdout clean for the ONE JSON line: anything libraries print to fd 1 meanwhile (NCCL's version banner,
    torchrun notices) is sent to stderr; emit() writes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


GUARD = None



def widened_block(capi):
    """SURVEY 8(f)2/3 next to the CPU port on the host cores (rank 0, bounded samples): the generalized R/T branch on a grid whose
    every column has a low-velocity zone, and the fast-marching travel times of example1's geometry (11 periods x 8 sources)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    from mctomo_b200 import synth
    cores = os.cpu_count() or 1
    out = {}
    # -- generalized R/T: 32 x 32 columns, 40 nodes deep, 11 frequencies, Rayleigh phase
    nx = 32
    grid = synth.make_grid(nx, nx, 40)
    pts, par = synth.generate_model(grid, 300, 1002)
    vp, vs, rho, sid = [np.zeros(grid.shape) for _ in range(3)] + [np.zeros(grid.shape, np.int32)]
    orc.kdtree_to_grid(pts, par, grid, grid.cover_box(), vp, vs, rho, sid)
    rng = np.random.default_rng(9)
    k0 = rng.integers(8, 20, size=(nx, nx))
    for i in range(nx):
        for j in range(nx):
            vs[i, j, k0[i, j]:k0[i, j] + 5] = vs[i, j, 0] * 0.85
    vp[:] = 1.73 * vs
    rho[:] = 1.74 * vp ** 0.25
    freqs = synth.example1_freqs()
    opts = capi.disp_opts(raylov=1, phaseGroup=0, nmodes=0)
    win = (1, nx, 1, nx)
    capi.set_grt(True, orc.GRT_PAR_LIKELIHOOD)
    try:
        capi.surf_dispersion(vp, vs, rho, grid, win, freqs, opts, check=False)
        t = time.perf_counter()
        pv, gv, ie, inval, rc = capi.surf_dispersion(vp, vs, rho, grid, win, freqs, opts, check=False)
        gpu_s = time.perf_counter() - t
        st = capi.grt_stats()
    finally:
        capi.set_grt(False)
    sample = [(i, j) for i in range(0, nx, 4) for j in range(0, nx, 4)][:4 * cores]

    def one(ij):
        n, (th, a, b, r) = orc.convert_column(vp[ij], vs[ij], rho[ij], grid.dz)
        return orc.grt_modes(th, a, b, r, freqs, modetype=1, phaseGroup=0, par=orc.GRT_PAR_LIKELIHOOD, math_mode=orc.LIBM)[0]
    t = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(one, sample))
    cpu_s = time.perf_counter() - t
    out["grt"] = {"what": "generalized R/T branch of surfmodes: every column of a 32x32x40 grid has a low-velocity zone, 11 frequencies, Rayleigh phase, "
                          "host arrays in / maps out (mct_surf_dispersion with mct_set_grt)",
                  "columns": st["columns"], "gpu_ms": 1e3 * gpu_s, "gpu_columns_per_s": st["columns"] / gpu_s, "ierr1_columns": int((ie == 1).sum()),
                  "cpu_columns_per_s": len(sample) / cpu_s, "cpu_sample_columns": len(sample), "cpu_cores": cores, "cpu_kind": "port (oracle/grt_ref.c, libm)"}
    # -- fm2d: example1's geometry
    n, nmaps, nsrc, nrc = 101, 11, 8, 8
    x, y = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n), indexing="ij")
    vel = np.zeros((nmaps, n + 2, n + 2))
    for m in range(nmaps):
        v = 2.5 + 0.1 * m + 0.2 * np.sin(5 * x + 3 * y + m) + 0.15 * np.cos(4 * y - 2 * x)
        vel[m, 1:-1, 1:-1] = v
        vel[m, 0, :] = vel[m, 1, :]; vel[m, -1, :] = vel[m, -2, :]; vel[m, :, 0] = vel[m, :, 1]; vel[m, :, -1] = vel[m, :, -2]
    src = rng.uniform(-4.5, 4.5, (nsrc, 2)); rcv = rng.uniform(-4.5, 4.5, (nrc, 2))
    srs = np.ones((nsrc, nrc), np.int32)
    o = capi.fm2d_opts()
    capi.fm2d_times(src, rcv, srs, vel, -5.0, -5.0, 0.1, 0.1, o)
    t = time.perf_counter()
    tt, _ = capi.fm2d_times(src, rcv, srs, vel, -5.0, -5.0, 0.1, 0.1, o)
    gpu_s = time.perf_counter() - t
    t = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        res = list(ex.map(lambda m: orc.fm2d_times(src, rcv, srs, vel[m], -5.0, -5.0, 0.1, 0.1)[1], range(nmaps)))
    cpu_s = time.perf_counter() - t
    out["fm2d"] = {"what": "fast-marching travel times of modrays (phase-velocity data): example1's geometry, 101x101 nodes, 11 periods x 8 sources = 88 "
                           "eikonal problems, shipped settings (mixed order, source refinement 4 x 8), host arrays in / times out (mct_fm2d_times)",
                   "problems": nmaps * nsrc, "gpu_ms": 1e3 * gpu_s, "cpu_ms": 1e3 * cpu_s, "cpu_threads": min(cores, nmaps), "cpu_kind": "port (oracle/fm2d_ref.c), one period per thread",
                   "bit_identical_to_port": bool(all(np.array_equal(res[m], tt[m]) for m in range(nmaps)))}
    return out


def main():
    global GUARD
    args = parse()
    GUARD = StdoutGuard()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from mctomo_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi.init(local)

    column_sharded = args.config == "C5" and world > 1
    grid, models, freqs, batch, spec = workload(args, 0 if column_sharded else rank, world)
    opts = capi.disp_opts(**spec)
    nm = max(spec["nmodes"], 1)
    nout = nm * len(freqs)
    pts, par, off = capi.pack_models(models)
    # x-slab of this rank (column sharding) or the whole grid
    if column_sharded:
        per = (grid.nx + world - 1) // world
        slab = (rank * per + 1, min(grid.nx, (rank + 1) * per))
    else:
        slab = (1, grid.nx)
    wx = slab[1] - slab[0] + 1
    ncell = grid.nx * grid.ny * grid.nz
    d_vp = torch.empty(batch * ncell, dtype=torch.float64, device=dev)
    d_vs = torch.empty_like(d_vp)
    d_rho = torch.empty_like(d_vp)
    d_sid = torch.empty(batch * ncell, dtype=torch.int32, device=dev)
    out_cols = per * world * grid.ny if column_sharded else batch * wx * grid.ny  # sharded: the FULL maps, gathered in place
    d_pv = torch.empty(out_cols * nout, dtype=torch.float64, device=dev)
    d_gv = torch.empty_like(d_pv)
    d_ie = torch.empty(out_cols, dtype=torch.int32, device=dev)
    d_fl = torch.zeros(2 * batch, dtype=torch.int32, device=dev)
    if column_sharded:
        capi.comm_init_torch(dist, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # a dedicated (non-default) stream: the library launches on the handle it is given, torch's events and
    # NCCL calls are recorded on the same one
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # config C3 = Rayleigh AND Love, phase + group, fundamental + first overtone: the Love maps are a second
    # dispersion pass over the model the first pass left resident (stage 1 runs once, as in the reference)
    love = args.config == "C3"
    love_opts = capi.disp_opts(raylov=0, phaseGroup=1, nmodes=2) if love else None
    if love:
        d_pv2 = torch.empty_like(d_pv); d_gv2 = torch.empty_like(d_pv); d_ie2 = torch.empty_like(d_ie)
        d_fl2 = torch.zeros(2 * batch, dtype=torch.int32, device=dev)

    def resident_step():
        if column_sharded:  # slab + in-place NCCL all-gather of the maps + flag all-reduce, all inside the library
            capi.forward_sharded_dev(grid, freqs, opts, d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), d_sid.data_ptr(),
                                     d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), stream)
            return
        capi.forward_batch_dev(grid, batch, freqs, opts, d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), d_sid.data_ptr(),
                               d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), stream, slab=slab)
        if love:
            capi.surf_dispersion_dev(d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), grid, (slab[0], slab[1], 1, grid.ny), freqs,
                                     love_opts, d_pv2.data_ptr(), d_gv2.data_ptr(), d_ie2.data_ptr(), d_fl2.data_ptr(), stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    capi.set_nuclei_batch(pts, par, off)
    solves_per_step_rank = batch * wx * grid.ny * nout * (2 if love else 1)
    if love:
        assert batch == 1
    # ---- warm-up
    for _ in range(args.warmup):
        resident_step()
    barrier()
    # ---- resident timing: EXACTLY K steps, CUDA events per step on the launching stream, L2 flushed between steps
    capi.reset_stats()
    capi.set_profiling(True)
    capi.kernel_times(reset=True)
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        resident_step()
        b.record()
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kt = capi.kernel_times(reset=True)
    st = capi.stats()
    main_launch = capi.last_launch()
    if love:  # roofline figures refer to the Rayleigh pass only: re-measure it alone (untimed for `value`)
        capi.reset_stats()
        for _ in range(args.steps):
            capi.surf_dispersion_dev(d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), grid, (slab[0], slab[1], 1, grid.ny), freqs,
                                     opts, d_pv.data_ptr(), d_gv.data_ptr(), d_ie.data_ptr(), d_fl.data_ptr(), stream)
        kt_all = kt
        kt = capi.kernel_times(reset=True)
        kt["k1_ms"] = kt_all["k1_ms"]
        st = dict(capi.stats(), n_launches=st["n_launches"])
    # latency of a proposal-sized call (20 x 20 columns of model 0, device-resident): what one rjMCMC step issues
    pw = (1, min(20, grid.nx), 1, min(20, grid.ny))
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_pp = torch.empty(400 * nout, dtype=torch.float64, device=dev); d_pg = torch.empty_like(d_pp)
    d_pi = torch.empty(400, dtype=torch.int32, device=dev); d_pf = torch.zeros(2, dtype=torch.int32, device=dev)
    for it in range(4):
        if it == 1:
            pa.record()
        capi.surf_dispersion_dev(d_vp.data_ptr(), d_vs.data_ptr(), d_rho.data_ptr(), grid, pw, freqs, opts, d_pp.data_ptr(),
                                 d_pg.data_ptr(), d_pi.data_ptr(), d_pf.data_ptr(), stream)
    pb.record()
    torch.cuda.synchronize()
    proposal_ms = pa.elapsed_time(pb) / 3
    capi.kernel_times(reset=True)
    capi.set_profiling(False)
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = solves_per_step_rank * world * args.steps / (total_ms * 1e-3)

    # ---- e2e: host nuclei -> C ABI -> host maps (pinned), every step
    e2e = None
    if not column_sharded and not love:
        h_pv = torch.empty((batch, grid.nx, grid.ny, nout), dtype=torch.float64).pin_memory()
        h_gv = torch.empty_like(h_pv).pin_memory()
        h_ie = torch.empty((batch, grid.nx, grid.ny), dtype=torch.int32).pin_memory()
        out = {"pvel": h_pv.numpy(), "gvel": h_gv.numpy(), "ierr": h_ie.numpy()}
        for _ in range(max(1, args.warmup - 1)):
            capi.forward_eval_batch(pts, par, off, grid, freqs, opts, out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = capi.forward_eval_batch(pts, par, off, grid, freqs, opts, out=out)  # synchronous: results are in host memory
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        assert int(r["ierr"].max()) <= 1
        e2e = {"value": solves_per_step_rank * world * args.steps / float(t_e2e.item()), "unit": "column*period solves/s",
               "h2d_bytes_per_step": int(pts.nbytes + par.nbytes + off.nbytes) * 2,
               "d2h_bytes_per_step": int(h_pv.numel() * 8 * 2 + h_ie.numel() * 4),
               "ms_per_step": float(t_e2e.item()) / args.steps * 1e3,
               "api": "mct_forward_eval_batch (host pointers in/out, tree build + H2D + kernels + D2H inside the timed region)"}

    c5 = None
    if not args.no_c5 and args.config == "C2":
        del d_vp, d_vs, d_rho, d_sid, d_pv, d_gv, d_ie
        torch.cuda.empty_cache()
        c5 = c5_block(args, torch, dist, capi, dev, rank, world, stream)
    if rank == 0:
        fl, fc = (F_LAYER_R, F_CALL_R) if spec["raylov"] == 1 else (F_LAYER_L, F_CALL_L)
        # EXECUTED work of all timed K2 launches on rank 0 (bit-identical columns are solved once: the represented
        # counts, equal to the reference's own, are reported under "work")
        w2 = st["n_dltar_executed"] * fc + st["n_layer_steps_executed"] * fl
        k2_s = kt["k2_ms"] * 1e-3
        probe = capi.fp64_peak_probe()
        achieved = w2 / k2_s / 1e12 if k2_s > 0 else None
        li = main_launch                                          # what the library itself says it launched
        sm_count = li["sm_count"]
        traffic, traffic_note = None, "no ncu capture of this build's K2 sources under profiles/k2_traffic.json"
        try:  # DRAM bytes per K2 launch from the ncu capture of THIS build: dropped when the K2 sources have changed since
            tj = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))
            if args.config == "C2" and batch == 32 and not column_sharded:
                if tj.get("k2_source_sha") == k2_source_sha() and tj.get("kernel") == li["kernel"]:
                    traffic, traffic_note = tj["dram_bytes_per_launch"], tj.get("source", "")
                else:
                    traffic_note = "profiles/k2_traffic.json was captured from different K2 sources or another kernel: dropped"
        except Exception:
            pass
        nominal = sm_count * 64 * 2 * (clocks.get("sm_max_mhz") or 0) * 1e6 / 1e12  # 64 FP64 FMA lanes per SM
        roof = {"bound": "fp64", "achieved": achieved, "peak": probe["dfma_tflops"], "unit": "TFLOP/s",
                "frac": achieved / probe["dfma_tflops"] if achieved else None, "traffic": traffic, "traffic_source": traffic_note,
                "kernel": li["kernel"], "kernel_columns": li["columns"], "kernel_columns_solved": li["columns_solved"],
                "kernel_lanes_per_column": li["lanes_per_column"],
                "k2_ms_per_step": kt["k2_ms"] / args.steps,
                "k2_share_of_step": kt["k2_ms"] / sum(step_ms),
                "peak_source": "measured live: 8 independent DFMA chains/thread (mct_fp64_peak_probe); MEASURED_PEAKS.json has no FP64 entry",
                "peak_nominal_tflops": nominal, "peak_nominal_what": f"{sm_count} SMs x 64 FP64 lanes x 2 flop x {clocks.get('sm_max_mhz')} MHz",
                "peak_dmul_dadd_tflops": probe["dmul_dadd_tflops"],
                "note": "achieved = nominal FP64 ops EXECUTED (184/layer step + 31/call, Rayleigh; SURVEY 8d) / K2 time; the kernel is "
                        "compiled without FMA contraction for bit-parity, so its ceiling is the DMUL+DADD rate",
                "layer_steps_per_s": st["n_layer_steps_executed"] / k2_s if k2_s > 0 else None,
                "k1_ms_per_step": kt["k1_ms"] / args.steps,
                "k1_hbm_gbs": (28.0 * batch * wx * grid.ny * grid.nz) * args.steps / (kt["k1_ms"] * 1e-3) / 1e9 if kt["k1_ms"] > 0 else None,
                "k1_hbm_peak_gbs": hbm_peak(), "k1_frac": ((28.0 * batch * wx * grid.ny * grid.nz) * args.steps / (kt["k1_ms"] * 1e-3) / 1e9 / hbm_peak())
                if kt["k1_ms"] > 0 else None}
        cpu = None
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            cspec = [spec, dict(raylov=0, phaseGroup=1, nmodes=2)] if love else spec
            r, e, sample, detail = cpu_sample(grid, models, freqs, cspec, args.cpu_seconds, cores)
            cpu = {"value": r, "unit": "column*period solves/s", "cores": cores, "kind": cpu_sample.kind, "sample": sample,
                   "forward_evals_per_sec": e, "detail": detail}
            if e2e:
                # share of the forward model's wall time that now runs on the GPU: what is left on the host in the
                # end-to-end call (tree build, staging, launches) against the CPU path's time for the same evaluations
                host_ms = max(0.0, e2e["ms_per_step"] - (kt["k1_ms"] + kt["k2_ms"] + kt["other_ms"]) / args.steps)
                cpu_ms = batch / e * 1e3
                cpu["gpu_share_of_forward_wall_time"] = 1.0 - host_ms / cpu_ms
                cpu["host_residual_ms_per_step"] = host_ms
        out = {"metric": "dispersion_solves_per_sec", "value": value, "unit": "column*period solves/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
               "scaling": "strong" if column_sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "forward_evals_per_sec": batch * world * args.steps / (total_ms * 1e-3) if not column_sharded else args.steps / (total_ms * 1e-3),
               "waves": "Rayleigh + Love" if love else "Rayleigh",
               # one solve = one (column, period, wave, mode) output; a group-velocity output costs two root searches
               "getsol_calls_per_sec": value * (2 if spec["phaseGroup"] else 1),
               "config": workload_config(args, grid, freqs, spec, len(models[0][0]), batch),
               "execution": {"parallelism": ("x-slab column sharding + in-place ncclAllGather of the maps inside the library (mct_forward_sharded_dev)" if column_sharded
                                             else f"chains sharded, {world} x {batch} independent models, no collective"),
                             "model_bytes_per_step": int(batch * ncell * 28)},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(st["n_launches"]), "roofline": roof, "cpu_baseline": cpu,
               "proposal_latency": {"columns": (pw[1] - pw[0] + 1) * (pw[3] - pw[2] + 1), "ms": proposal_ms,
                                    "what": "check_model + layering + dispersion of a 20x20-column window, device-resident model, cooperative kernel (several warps per column)"},
               "c5": c5,
               "work": {"dltar_calls_per_step": st["n_dltar"] / args.steps, "layer_steps_per_step": st["n_layer_steps"] / args.steps,
                        "columns_per_step": st["n_columns"] / args.steps,
                        "executed": {"dltar_calls_per_step": st["n_dltar_executed"] / args.steps,
                                     "layer_steps_per_step": st["n_layer_steps_executed"] / args.steps,
                                     "distinct_columns_per_step": st["n_columns_solved"] / args.steps},
                        "what": "represented = the reference's own call counts for these inputs (identical to the oracle's); "
                                "executed = after folding bit-identical columns"}}
        if world == 1 and not args.no_widened and not column_sharded:
            try:
                out["widened"] = widened_block(capi)
            except Exception as e:  # never lose the headline line over the side block
                out["widened"] = {"error": repr(e)}
        GUARD.emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    capi.shutdown()


if __name__ == "__main__":
    main()
