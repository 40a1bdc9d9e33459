// k2_dedup.cuh -- exact column de-duplication in front of the dispersion kernels.
//
// A Voronoi model has far fewer distinct velocity columns than grid columns: neighbouring columns cut the same cells at
// the same depths and become, after convert_to_layer (reference src/likelihood_surf.F90:523-629) and the narrowing to
// real*4 (surfmodes/surfmodes.f90:81-83), bit-identical layer stacks.  surfdisp96 is a pure function of that stack,
// so such columns are solved ONCE and their outputs copied -- an exact optimisation: the
// outputs, ierr included, are the representative's bits.  Measured distinct fractions on the bench's model family:
// C2 0.97, C3 0.85, C5 slab 0.58, C4 0.33-0.85 (25-300 cells).
//
// Three small launches over the columns (layer records are read once, coalesced):
//   1. hash the stack (layer count + every float4 record, bitwise; the model index is part of the key, so columns of
//      different chains are never merged -- a rejected model's columns must stay unsolved) and insert the column into
//      an open-addressing table (64-bit slots: hash tag | column + 1); a hit is confirmed by comparing the two stacks
//      record for record, so a hash collision can only cost time, never correctness;
//   2. per group: the smallest column index (deterministic representative, whichever thread won the insert) and size;
//   3. per column: representative, multiplicity, the status K2 sees (< 0 for a duplicate) and its sort key (0 for a
//      duplicate: the counting sort then parks all duplicates behind the columns that are really solved).
// After K2, dedup_scatter_kernel copies pvel/gvel/ierr from representative to duplicates.
#pragma once

#define MCT_ST_DUP (-1)

__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33;
  return h;
}

__device__ __forceinline__ bool same_stack(const float4* __restrict__ lay, int stride, int a, int b, int n) {
  for (int m = 0; m < n; ++m) {
    const float4 x = lay[(size_t)m * stride + a], y = lay[(size_t)m * stride + b];
    if (__float_as_uint(x.x) != __float_as_uint(y.x) || __float_as_uint(x.y) != __float_as_uint(y.y) ||
        __float_as_uint(x.z) != __float_as_uint(y.z) || __float_as_uint(x.w) != __float_as_uint(y.w))
      return false;
  }
  return true;
}

// table: nslots (power of two) 64-bit slots, zero = empty.  rep0[c] = the column c was matched with (itself when new
// or not solvable).
__global__ void __launch_bounds__(128) dedup_insert_kernel(const float4* __restrict__ lay, const int32_t* __restrict__ nlay,
                                                           const int32_t* __restrict__ status, int ncol, int stride, int cols_per_model,
                                                           unsigned long long* table, unsigned nslots, int32_t* __restrict__ rep0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  if (status[c] != 0) { rep0[c] = c; return; } // columns that are not solved are not folded
  const int n = nlay[c];
  const int model = c / cols_per_model;
  unsigned long long h = mix64(0x1234567ull + (unsigned long long)model, (unsigned long long)n);
  for (int m = 0; m < n; ++m) {
    const float4 L = lay[(size_t)m * stride + c];
    h = mix64(h, ((unsigned long long)__float_as_uint(L.x) << 32) | __float_as_uint(L.y));
    h = mix64(h, ((unsigned long long)__float_as_uint(L.z) << 32) | __float_as_uint(L.w));
  }
  const unsigned long long tag = h & 0xffffffff00000000ull;
  const unsigned long long mine = tag | (unsigned long long)(unsigned)(c + 1);
  unsigned slot = (unsigned)(h * 0x9e3779b97f4a7c15ull >> 32) & (nslots - 1);
  for (unsigned probe = 0; probe < nslots; ++probe) {
    unsigned long long cur = table[slot];
    if (cur == 0ull) {
      cur = atomicCAS(&table[slot], 0ull, mine);
      if (cur == 0ull) { rep0[c] = c; return; }
    }
    if ((cur & 0xffffffff00000000ull) == tag) {
      const int o = (int)(unsigned)(cur & 0xffffffffull) - 1;
      if (o / cols_per_model == model && nlay[o] == n && same_stack(lay, stride, c, o, n)) { rep0[c] = o; return; }
    }
    slot = (slot + 1) & (nslots - 1);
  }
  rep0[c] = c; // table full (cannot happen: nslots >= 2 ncol): solve the column itself
}

__global__ void __launch_bounds__(256) dedup_group_kernel(const int32_t* __restrict__ rep0, int ncol, int32_t* __restrict__ minrep,
                                                          int32_t* __restrict__ mult0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  const int r = rep0[c];
  atomicMin(&minrep[r], c);
  atomicAdd(&mult0[r], 1);
}

// rep[c]: final representative; mult[c] (representatives only): group size; kstat / skey: what K2 and the sort see;
// neff: number of columns K2 has to visit (representatives + unsolvable columns).
__global__ void __launch_bounds__(256) dedup_finish_kernel(const int32_t* __restrict__ rep0, const int32_t* __restrict__ minrep,
                                                           const int32_t* __restrict__ mult0, const int32_t* __restrict__ status,
                                                           const int32_t* __restrict__ nlay, int ncol, int32_t* __restrict__ rep,
                                                           int32_t* __restrict__ mult, int32_t* __restrict__ kstat, int32_t* __restrict__ skey,
                                                           int32_t* neff, const float4* __restrict__ lay, int stride, float qscale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  bool mine = false;
  if (c < ncol) {
    const int g = rep0[c];
    const int r = minrep[g];
    rep[c] = r;
    if (r == c) {
      mult[c] = mult0[g];
      kstat[c] = status[c];
      int key = max(nlay[c], 1);
      if (qscale > 0.f) {
        // secondary key inside a layer-count bin: (bottom vs - 0.79 top vs) / dc is roughly how far getsol's scan walks
        // over the band, i.e. a proxy of the column's number of secular-function evaluations.  Whole km/s (qscale = 1)
        // splits a bin in two or three, longer searches first: 1.7-3.2 % on C1/C2/C4, nothing on C3/C5; finer levels
        // (2..6 per km/s) give the gain back by scattering spatial neighbours (tools/k2_ab.py, MCT_SORT_PROXY)
        const int n = key;
        const float btop = lay[c].z > 0.01f ? lay[c].z : lay[(size_t)(n > 1 ? 1 : 0) * stride + c].z;
        const float proxy = lay[(size_t)(n - 1) * stride + c].z - 0.786f * btop;
        const int q = min(7, max(0, (int)(proxy * qscale)));
        key = min(key, 31) * 8 + q;
      }
      skey[c] = key;
      mine = true;
    } else {
      mult[c] = 0;
      kstat[c] = MCT_ST_DUP;
      skey[c] = 0;
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(neff, __popc(b));
}

// outputs of the duplicates: one thread per (column, output index)
__global__ void __launch_bounds__(256) dedup_scatter_kernel(const int32_t* __restrict__ rep, int ncol, int nout, const int32_t* __restrict__ skip,
                                                            int cols_per_model, double* __restrict__ pvel, double* __restrict__ gvel,
                                                            int32_t* __restrict__ ierr) {
  const long long n = (long long)ncol * nout;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t / nout);
    const int k = (int)(t - (long long)c * nout);
    const int r = rep[c];
    if (r == c) continue;
    if (skip && skip[2 * (c / cols_per_model)] != 0) continue; // model rejected by check_model: nothing is written
    pvel[(size_t)c * nout + k] = pvel[(size_t)r * nout + k];
    gvel[(size_t)c * nout + k] = gvel[(size_t)r * nout + k];
    if (k == 0) ierr[c] = ierr[r];
  }
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t* p, long long n, int32_t v) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) p[t] = v;
}

// ---- compact inputs: the columns to solve, gathered in sorted order ---------------------------------------------------------
// K2 reads a column's layer records once per secular-function evaluation.  In the per-column layout the 32 columns of a warp
// are neighbours in the SORTED order but lie anywhere in memory (32 separate sectors per load, a footprint of the whole
// grid -- 4 GB of layer records at C5 against a 126 MB L2); gathered into list order a warp reads contiguous rows and the
// footprint shrinks to the distinct columns.  One thread per list position; duplicates are not copied.
__global__ void __launch_bounds__(128) compact_layers_kernel(const float4* __restrict__ lay, const int32_t* __restrict__ nlay,
                                                             const int32_t* __restrict__ status, const int32_t* __restrict__ list, int n,
                                                             int stride, int stride_c, float4* __restrict__ layc, int32_t* __restrict__ nlayc,
                                                             int32_t* __restrict__ statc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int col = list[t];
  const int nl = nlay[col];
  nlayc[t] = nl;
  statc[t] = status[col];
  for (int m = 0; m < nl; ++m) layc[(size_t)m * stride_c + t] = lay[(size_t)m * stride + col];
}

// ---- balanced sharding of one chain's columns over several ranks (mct_comm.cuh, mode 1) --------------------------------
// Every rank holds the same sorted list of columns to solve (perm[0..neff)).  It is dealt in CHUNKS of 32 consecutive
// entries (chunk c goes to rank c % n): a warp of the dispersion kernel then still holds 32 neighbours of the sorted
// order -- same layer count, adjacent columns, coalesced layer records.  (Dealing single entries was measured first:
// 1 061 ms per rank against 861 ms for the same number of columns, the warps' columns being n apart.)
#define SHARD_CHUNK 32
__host__ __device__ __forceinline__ long long shard_global(int i, int r, int n) { // local entry i of rank r -> list position
  return ((long long)(i / SHARD_CHUNK) * n + r) * SHARD_CHUNK + (i % SHARD_CHUNK);
}
__global__ void __launch_bounds__(256) shard_pick_kernel(const int32_t* __restrict__ perm, int neff, int r, int n, int nmine, int32_t* __restrict__ mine) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nmine) return;
  mine[i] = perm[shard_global(i, r, n)];
}
// this rank's results -> its chunk of the exchange buffers (per entries per rank)
__global__ void __launch_bounds__(256) shard_pack_kernel(const int32_t* __restrict__ perm, int neff, int r, int n, int per, int nout,
                                                         const double* __restrict__ pvel, const double* __restrict__ gvel,
                                                         const int32_t* __restrict__ ierr, double* __restrict__ bp, double* __restrict__ bg,
                                                         int32_t* __restrict__ bi) {
  const long long total = (long long)per * nout;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / nout), k = (int)(t - (long long)i * nout);
    const long long j = shard_global(i, r, n);
    if (j >= neff) continue;
    const int col = perm[j];
    const size_t o = ((size_t)r * per + i) * nout + k;
    bp[o] = pvel[(size_t)col * nout + k];
    if (bg) bg[o] = gvel[(size_t)col * nout + k];
    if (k == 0) bi[(size_t)r * per + i] = ierr[col];
  }
}
// the other ranks' results -> the maps
__global__ void __launch_bounds__(256) shard_unpack_kernel(const int32_t* __restrict__ perm, int neff, int r, int n, int per, int nout,
                                                           double* __restrict__ pvel, double* __restrict__ gvel, int32_t* __restrict__ ierr,
                                                           const double* __restrict__ bp, const double* __restrict__ bg,
                                                           const int32_t* __restrict__ bi) {
  const long long total = (long long)neff * nout;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / nout), k = (int)(t - (long long)j * nout);
    const int c = j / SHARD_CHUNK;
    const int rk = c % n, i = (c / n) * SHARD_CHUNK + (j % SHARD_CHUNK);
    if (rk == r) continue;
    const int col = perm[j];
    const size_t o = ((size_t)rk * per + i) * nout + k;
    pvel[(size_t)col * nout + k] = bp[o];
    if (bg) gvel[(size_t)col * nout + k] = bg[o];
    if (k == 0) ierr[col] = bi[(size_t)rk * per + i];
  }
}
