// mct_comm.cuh -- the multi-GPU data plane behind the C ABI (included by mct_api.cu).
//
// MCTomo's parallel model is one MPI rank per chain (reference src/MCTomo.F90:82-86,131-133): chains never talk
// on the data path.  The ONE configuration with an exchange step is a single chain whose grid is too large for one
// GPU's patience (BASELINE config 5): the columns are cut into contiguous x-slabs, one per rank, every rank grids
// and solves its slab, and the per-period dispersion maps are all-gathered so every rank (and its host, which runs
// fm2d on the whole map, src/likelihood_surf.F90:295-336) holds the full (np,ny,nx) field; check_model's whole-grid
// `any` (:631-646) becomes a MAX all-reduce of one flag.  SURVEY.md section 8(e).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library keeps working, single-GPU, on a box without NCCL,
// and inside a process that already loaded a NCCL (PyTorch's bundled one) it binds to that very copy instead of a
// second one.  The host program -- Fortran + MPI in the reference -- only has to move the 128-byte unique id from
// rank 0 to the others (MPI_Bcast; torch.distributed.broadcast in bench.py) and call mct_comm_init.
//
// The all-gather is IN PLACE: the map buffer (nout, ny, per*nranks) is both send and receive buffer, rank r's slab
// being the r-th chunk of it, which is exactly where the slab-restricted forward evaluation writes its outputs.
#pragma once
#include <dlfcn.h>
#include <nccl.h> // types and enums only; every function is resolved with dlsym

namespace {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

struct Comm {
  int mode = 0; // 0: contiguous x-slabs; 1: balanced -- every rank grids the whole model, solves every n-th
                // representative of the sorted list, the compact results are all-gathered
  NcclApi api;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int version = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr}; // brackets the last all-gather (mct_comm_last_ms)
  bool timed = false;
};
Comm gc;

int nccl_load() {
  if (gc.api.h) return MCT_OK;
  const char* names[] = {getenv("MCT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(MCT_E_CUDA, "comm: libnccl.so.2 not found (%s); set MCT_NCCL_LIB", dlerror());
  NcclApi a;
  a.h = h;
#define MCT_SYM(field, name)                                                        \
  *(void**)(&a.field) = dlsym(h, name);                                             \
  if (!a.field) { dlclose(h); return fail(MCT_E_CUDA, "comm: %s missing from the NCCL library", name); }
  MCT_SYM(GetUniqueId, "ncclGetUniqueId")
  MCT_SYM(CommInitRank, "ncclCommInitRank")
  MCT_SYM(CommDestroy, "ncclCommDestroy")
  MCT_SYM(AllGather, "ncclAllGather")
  MCT_SYM(AllReduce, "ncclAllReduce")
  MCT_SYM(GetErrorString, "ncclGetErrorString")
  MCT_SYM(GetVersion, "ncclGetVersion")
#undef MCT_SYM
  gc.api = a;
  gc.api.GetVersion(&gc.version);
  return MCT_OK;
}

#define NK(call)                                                                                              \
  do {                                                                                                        \
    ncclResult_t r_ = (call);                                                                                 \
    if (r_ != ncclSuccess) return fail(MCT_E_CUDA, "%s failed: %s", #call, gc.api.GetErrorString(r_));        \
  } while (0)

// balanced mode: this rank's results -> exchange buffers, in-place all-gather, the others' results -> maps.  Called by
// launch_k2 right after the dispersion kernel, before the duplicates are filled from their representatives.
int shard_exchange(int neff, int nout, bool with_group, double* d_pvel, double* d_gvel, int32_t* d_ierr, cudaStream_t st) {
  const int n = g.shard_n, r = g.shard_r;
  const int nchunks = (neff + SHARD_CHUNK - 1) / SHARD_CHUNK;
  const int per = ((nchunks + n - 1) / n) * SHARD_CHUNK; // list entries per rank (whole chunks)
  if (per == 0) return MCT_OK;
  int rc;
  const size_t np = (size_t)n * per * nout;
  if ((rc = ensure(g.sh_p, np * 8)) || (rc = ensure(g.sh_i, (size_t)n * per * 4 + 8))) return rc;
  if (with_group && (rc = ensure(g.sh_g, np * 8))) return rc;
  double* bp = (double*)g.sh_p.p;
  double* bg = with_group ? (double*)g.sh_g.p : nullptr;
  int32_t* bi = (int32_t*)g.sh_i.p;
  {
    ProfScope ps(2, st);
    shard_pack_kernel<<<grid_blocks((long long)per * nout, 256, 16), 256, 0, st>>>((const int32_t*)g.perm.p, neff, r, n, per, nout, d_pvel,
                                                                                  d_gvel, d_ierr, bp, bg, bi);
  }
  CK(cudaGetLastError());
  if (gc.ev[0]) CK(cudaEventRecord(gc.ev[0], st));
  const int64_t perb = (int64_t)per * nout * 8;
  if ((rc = mct_allgather_inplace(bp, perb, st))) return rc;
  if (bg && (rc = mct_allgather_inplace(bg, perb, st))) return rc;
  if ((rc = mct_allgather_inplace(bi, (int64_t)per * 4, st))) return rc;
  if (gc.ev[1]) { CK(cudaEventRecord(gc.ev[1], st)); gc.timed = true; }
  {
    ProfScope ps(2, st);
    shard_unpack_kernel<<<grid_blocks((long long)neff * nout, 256, 16), 256, 0, st>>>((const int32_t*)g.perm.p, neff, r, n, per, nout, d_pvel,
                                                                                     d_gvel, d_ierr, bp, bg, bi);
  }
  CK(cudaGetLastError());
  g.host_stats.n_launches += 2;
  return MCT_OK;
}

void comm_release() {
  if (gc.comm && gc.api.CommDestroy) gc.api.CommDestroy(gc.comm);
  gc.comm = nullptr;
  gc.rank = 0;
  gc.nranks = 1;
  for (auto& e : gc.ev) { if (e) cudaEventDestroy(e); e = nullptr; }
  gc.timed = false;
}

} // namespace

extern "C" {

int mct_comm_unique_id(void* id128) {
  if (!id128) return fail(MCT_E_INVALID_ARG, "comm_unique_id: NULL pointer");
  int rc = nccl_load();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == MCT_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  NK(gc.api.GetUniqueId(&u));
  memcpy(id128, &u, sizeof u);
  return MCT_OK;
}

int mct_comm_init(const void* id128, int rank, int nranks) {
  NEED_INIT();
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(MCT_E_INVALID_ARG, "comm_init: bad arguments");
  if (gc.comm) return fail(MCT_E_INVALID_ARG, "comm_init: already initialised (rank %d of %d); call mct_comm_destroy first", gc.rank, gc.nranks);
  int rc = nccl_load();
  if (rc) return rc;
  ncclUniqueId u;
  memcpy(&u, id128, sizeof u);
  CK(cudaSetDevice(g.device));
  NK(gc.api.CommInitRank(&gc.comm, nranks, u, rank));
  gc.rank = rank;
  gc.nranks = nranks;
  CK(cudaEventCreate(&gc.ev[0]));
  CK(cudaEventCreate(&gc.ev[1]));
  return MCT_OK;
}

int mct_comm_destroy(void) {
  comm_release();
  return MCT_OK;
}

int mct_comm_set_mode(int mode) {
  if (mode != 0 && mode != 1) return fail(MCT_E_INVALID_ARG, "comm_set_mode: 0 (contiguous x-slabs) or 1 (balanced)");
  gc.mode = mode;
  return MCT_OK;
}

int mct_comm_info(int* rank, int* nranks, int* nccl_version) {
  if (rank) *rank = gc.rank;
  if (nranks) *nranks = gc.nranks;
  if (nccl_version) *nccl_version = gc.version;
  return gc.comm ? MCT_OK : MCT_E_INVALID_ARG;
}

// x-slab of a rank: contiguous, equal width per = ceil(nx/nranks); the last slabs may be short or empty (ix1 < ix0)
int mct_slab_bounds(int nx, int nranks, int rank, int* ix0, int* ix1, int* per) {
  if (nx < 1 || nranks < 1 || rank < 0 || rank >= nranks || !ix0 || !ix1) return fail(MCT_E_INVALID_ARG, "slab_bounds: bad arguments");
  const int p = (nx + nranks - 1) / nranks;
  *ix0 = rank * p + 1;
  *ix1 = std::min(nx, (rank + 1) * p);
  if (per) *per = p;
  return MCT_OK;
}

// In-place all-gather of a device buffer made of nranks chunks of bytes_per_rank bytes; this rank's chunk is the
// rank-th.  One rank: nothing to do.
int mct_allgather_inplace(void* d_buf, int64_t bytes_per_rank, void* stream) {
  NEED_INIT();
  if (!d_buf || bytes_per_rank < 0) return fail(MCT_E_INVALID_ARG, "allgather: bad arguments");
  if (gc.nranks == 1 || bytes_per_rank == 0) return MCT_OK;
  if (!gc.comm) return fail(MCT_E_INVALID_ARG, "allgather: mct_comm_init has not been called");
  cudaStream_t st = pick(stream);
  char* base = (char*)d_buf;
  // 8-byte elements when possible (every map is f64 or an even number of int32 per column row in practice)
  if (bytes_per_rank % 8 == 0 && ((uintptr_t)base % 8) == 0)
    NK(gc.api.AllGather(base + (size_t)gc.rank * bytes_per_rank, base, (size_t)bytes_per_rank / 8, ncclFloat64, gc.comm, st));
  else
    NK(gc.api.AllGather(base + (size_t)gc.rank * bytes_per_rank, base, (size_t)bytes_per_rank, ncclInt8, gc.comm, st));
  return MCT_OK;
}

// MAX all-reduce of n int32 status flags, in place ({model_invalid, max condition code}: check_model is an `any`
// over the whole grid, src/likelihood_surf.F90:631-646).
int mct_allreduce_flags(int32_t* d_flags, int n, void* stream) {
  NEED_INIT();
  if (!d_flags || n < 1) return fail(MCT_E_INVALID_ARG, "allreduce_flags: bad arguments");
  if (gc.nranks == 1) return MCT_OK;
  if (!gc.comm) return fail(MCT_E_INVALID_ARG, "allreduce_flags: mct_comm_init has not been called");
  NK(gc.api.AllReduce(d_flags, d_flags, (size_t)n, ncclInt32, ncclMax, gc.comm, pick(stream)));
  return MCT_OK;
}

// One chain's forward evaluation with its columns sharded over the communicator's ranks (config 5):
//   this rank grids and solves its x-slab (mct_slab_bounds) of the resident nuclei set (mct_set_nuclei_batch with
//   nb = 1, the same nuclei on every rank), writing straight into its chunk of the FULL maps
//       d_pvel, d_gvel : (nout, ny, per*nranks) doubles      d_ierr : (ny, per*nranks) int32
//   then the maps are all-gathered in place (pvel, ierr; gvel only when opt->phaseGroup == 1) and the two status
//   flags MAX-reduced, all on `stream`.  d_vp/d_vs/d_rho/d_sites are whole-grid arrays of which only the slab is
//   touched.  Columns beyond nx (padding of the last slab) are never written.
int mct_forward_sharded_dev(const mct_grid* gr, int derive_vp_rho, const double* freqs, int np, const mct_disp_opts* opt,
                            double* d_vp, double* d_vs, double* d_rho, int32_t* d_sites_id, double* d_pvel, double* d_gvel,
                            int32_t* d_ierr, int32_t* d_flags, void* stream) {
  NEED_INIT();
  if (!grid_ok(gr) || !opt || !d_pvel || !d_gvel || !d_ierr || !d_flags) return fail(MCT_E_INVALID_ARG, "forward_sharded: bad arguments");
  if (gc.nranks > 1 && !gc.comm) return fail(MCT_E_INVALID_ARG, "forward_sharded: mct_comm_init has not been called");
  if (gc.mode == 1 && gc.nranks > 1) {
    // balanced: K1 + maps + check + layering + de-duplication of the WHOLE grid on every rank (milliseconds), then every
    // rank solves every nranks-th entry of the sorted list of distinct columns; compact results all-gathered in place
    // (launch_k2 / shard_exchange).  Outputs are the plain (nout, ny, nx) maps, identical on every rank; the flags need
    // no reduction (every rank checked the whole grid).
    g.shard_n = gc.nranks;
    g.shard_r = gc.rank;
    gc.timed = false;
    const int rcb = mct_forward_batch_dev(gr, 1, derive_vp_rho, 1, gr->nx, freqs, np, opt, d_vp, d_vs, d_rho, d_sites_id, d_pvel, d_gvel,
                                          d_ierr, d_flags, stream);
    g.shard_n = 1;
    g.shard_r = 0;
    return rcb;
  }
  int ix0, ix1, per;
  int rc = mct_slab_bounds(gr->nx, gc.nranks, gc.rank, &ix0, &ix1, &per);
  if (rc) return rc;
  const int nm = opt->nmodes <= 0 ? 1 : opt->nmodes;
  const size_t nout = (size_t)np * nm;
  const size_t col0 = (size_t)gc.rank * per * gr->ny; // first column of this rank's chunk
  cudaStream_t st = pick(stream);
  if (ix1 >= ix0) {
    rc = mct_forward_batch_dev(gr, 1, derive_vp_rho, ix0, ix1, freqs, np, opt, d_vp, d_vs, d_rho, d_sites_id, d_pvel + col0 * nout,
                               d_gvel + col0 * nout, d_ierr + col0, d_flags, stream);
    if (rc) return rc;
  } else {
    CK(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int32_t), st)); // an empty slab has nothing to report
  }
  if (gc.nranks == 1) return MCT_OK;
  const int64_t cols = (int64_t)per * gr->ny;
  if (gc.ev[0]) CK(cudaEventRecord(gc.ev[0], st));
  if ((rc = mct_allgather_inplace(d_pvel, cols * (int64_t)nout * 8, stream))) return rc;
  if (opt->phaseGroup == 1 && (rc = mct_allgather_inplace(d_gvel, cols * (int64_t)nout * 8, stream))) return rc;
  if ((rc = mct_allgather_inplace(d_ierr, cols * 4, stream))) return rc;
  if ((rc = mct_allreduce_flags(d_flags, 2, stream))) return rc;
  if (gc.ev[1]) { CK(cudaEventRecord(gc.ev[1], st)); gc.timed = true; }
  return MCT_OK;
}

// Device time of the collectives of the last mct_forward_sharded_dev call (synchronises on its end event).
int mct_comm_last_ms(double* ms) {
  if (!ms) return fail(MCT_E_INVALID_ARG, "comm_last_ms: NULL pointer");
  *ms = 0.0;
  if (!gc.timed) return MCT_OK;
  CK(cudaEventSynchronize(gc.ev[1]));
  float t = 0.f;
  CK(cudaEventElapsedTime(&t, gc.ev[0], gc.ev[1]));
  *ms = (double)t;
  return MCT_OK;
}

} // extern "C"
