// kdtree_build.h -- host-side construction of the nearest-nucleus search tree that K1 walks.
//
// The tree has to be the one kdtree2_create builds (reference src/kdtree2.f90:609-689, 691-840,
// 842-901, 936-983): leaves of at most 13 points, split on the widest dimension of the node's
// (partly inherited) box at the arithmetic mean, `<= mean` to the left, cut_val_left/right from
// the children's exact boxes -- because which nucleus wins an exact distance tie depends on the
// traversal order this geometry induces.  n is at most a few thousand (48 B per nucleus), so the
// build is microseconds of host work per call; all per-node queries run on the GPU.
#pragma once
#include <cstdint>
#include <vector>

#include "k1_voronoi.cuh"

struct HostKdTree {
  std::vector<KdNodeDev> nodes;
  std::vector<int32_t> ind;  // position -> original 1-based index
  std::vector<double> rpts;  // rearranged coordinates (3, n)
  int root = -1;
  bool degenerate = false;   // the Fortran build would never terminate (> 13 points coincident in all dimensions)
};

class KdBuilder {
 public:
  KdBuilder(const double* pts, int n, HostKdTree& out) : p_(pts), n_(n), t_(out) {}
  void run() {
    t_.nodes.clear();
    t_.nodes.reserve(2 * (size_t)(n_ / 6 + 2));
    t_.ind.resize(n_);
    for (int j = 0; j < n_; ++j) t_.ind[j] = j + 1;
    t_.degenerate = false;
    t_.root = range(0, n_ - 1, -1, 0);
    t_.rpts.resize(3 * (size_t)n_);
    for (int i = 0; i < n_; ++i)
      for (int d = 0; d < 3; ++d) t_.rpts[3 * (size_t)i + d] = p_[3 * (size_t)(t_.ind[i] - 1) + d];
  }

 private:
  static constexpr int kBucket = 12;
  static constexpr int kMaxChain = 32;
  double coord(int d, int pos) const { return p_[3 * (size_t)(t_.ind[pos] - 1) + d]; }

  // min/max of coordinate d over positions l..u, scanning in pairs like spread_in_coordinate
  void extent(int d, int l, int u, double& lo, double& up) const {
    double smin = coord(d, l), smax = smin;
    int i = l + 2;
    for (; i <= u; i += 2) {
      double a = coord(d, i - 1), b = coord(d, i);
      if (a > b) { double t = a; a = b; b = t; }
      if (smin > a) smin = a;
      if (smax < b) smax = b;
    }
    if (i == u + 1) {
      double last = coord(d, u);
      if (smin > last) smin = last;
      if (smax < last) smax = last;
    }
    lo = smin;
    up = smax;
  }

  // Hoare-style partition around a VALUE (select_on_coordinate_value); returns last position <= alpha
  int partition(int d, double alpha, int l, int u) {
    int lb = l, rb = u;
    while (lb < rb) {
      if (coord(d, lb) <= alpha) ++lb;
      else { std::swap(t_.ind[lb], t_.ind[rb]); --rb; }
    }
    return (coord(d, lb) <= alpha) ? lb : lb - 1;
  }

  // chain = number of consecutive ancestors whose split left one child empty (same l..u handed down)
  int range(int l, int u, int parent, int chain) {
    if (u < l) return -1;
    const int id = (int)t_.nodes.size();
    t_.nodes.emplace_back();
    {
      KdNodeDev& N = t_.nodes[id];
      N.left = N.right = -1;
      N.l = l; N.u = u;
      N.cut_dim = -1;
      N.cut_val = N.cut_left = N.cut_right = 0.0;
      N.pad = 0;
    }
    if (u - l <= kBucket) {
      for (int d = 0; d < 3; ++d) extent(d, l, u, t_.nodes[id].lo[d], t_.nodes[id].up[d]);
      return id;
    }
    for (int d = 0; d < 3; ++d) {
      // only the dimension the parent cut along is re-measured; the others are inherited (:768-780)
      if (parent < 0 || d == t_.nodes[parent].cut_dim) extent(d, l, u, t_.nodes[id].lo[d], t_.nodes[id].up[d]);
      else { t_.nodes[id].lo[d] = t_.nodes[parent].lo[d]; t_.nodes[id].up[d] = t_.nodes[parent].up[d]; }
    }
    int c = 0;
    double widest = t_.nodes[id].up[0] - t_.nodes[id].lo[0];
    for (int d = 1; d < 3; ++d) {
      const double e = t_.nodes[id].up[d] - t_.nodes[id].lo[d];
      if (e > widest) { widest = e; c = d; } // maxloc: first maximum
    }
    double sum = 0.0;
    for (int i = l; i <= u; ++i) sum += coord(c, i);
    const double mean = sum / (double)(u - l + 1);
    const int m = partition(c, mean, l, u);
    // A split that leaves one side empty (every coordinate of the node equal along the chosen -- inherited, hence
    // over-estimated -- box dimension) is legal in kdtree2: the node keeps its single child, takes over the child's
    // box (:818-826), and the child re-measures that dimension and cuts along another one.  The search treats such
    // a node as terminal and scans its whole range l..u (:1388).  Only when the points of the range coincide in
    // every dimension does the Fortran recurse for ever; a chain of kMaxChain one-child levels is reported instead.
    const bool one_child = (m >= u || m < l);
    if (one_child && chain >= kMaxChain) { t_.degenerate = true; return id; }
    t_.nodes[id].cut_dim = c;
    const int left = range(l, m, id, one_child ? chain + 1 : 0);
    const int right = t_.degenerate ? -1 : range(m + 1, u, id, one_child ? chain + 1 : 0);
    if (t_.degenerate) return id;
    KdNodeDev& N = t_.nodes[id];
    N.left = left;
    N.right = right;
    if (right < 0 || left < 0) {
      const KdNodeDev& A = t_.nodes[right < 0 ? left : right];
      for (int d = 0; d < 3; ++d) { N.lo[d] = A.lo[d]; N.up[d] = A.up[d]; }
      if (right < 0) { N.cut_left = A.up[c]; N.cut_val = N.cut_left; }
      else { N.cut_right = A.lo[c]; N.cut_val = N.cut_right; }
      return id;
    }
    const KdNodeDev& A = t_.nodes[left];
    const KdNodeDev& B = t_.nodes[right];
    N.cut_right = B.lo[c];
    N.cut_left = A.up[c];
    N.cut_val = (N.cut_left + N.cut_right) / 2;
    for (int d = 0; d < 3; ++d) {
      N.up[d] = A.up[d] > B.up[d] ? A.up[d] : B.up[d];
      N.lo[d] = A.lo[d] < B.lo[d] ? A.lo[d] : B.lo[d];
    }
    return id;
  }

  const double* p_;
  int n_;
  HostKdTree& t_;
};
