// Native (host, OpenMP) bi-stride hierarchy builder: the integer part of the reference's
// BistrideMultiLayerGraph (src/graph_wrappers/bsms_graph_wrapper.py:58-154, graph_wrapper.py:67-134) —
// connected clusters, BFS distance parity from one seed per cluster, "keep the smaller of the even / odd sets",
// and the new adjacency = pattern of (A+I)^2 without its diagonal restricted to the kept nodes and re-indexed.
// The reference does this in pure Python (+ one MKL SpGEMM): 96 s / 5.5 GB for a 2 M-node mesh.
// Seeds (node nearest the cluster centroid, floating point) are chosen by the caller (hierarchy.py, numpy, the
// arithmetic already pinned to the reference's goldens); everything here is integer work and must be EXACT.
//
// Output edge order: row-major over kept nodes with SORTED columns (the reference inherits whatever order its
// SpGEMM leaves inside a row; tests compare edge sets).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/bsms_b200.h"

namespace {
struct Csr {
  std::vector<int64_t> ptr;
  std::vector<int32_t> col;
};
// CSR of the directed pattern row = g[0][e] -> col = g[1][e], duplicates removed, columns sorted
Csr build_csr(const int64_t* g, int64_t E, int64_t n) {
  Csr a;
  a.ptr.assign(n + 1, 0);
  for (int64_t e = 0; e < E; ++e) a.ptr[g[e] + 1]++;
  for (int64_t i = 0; i < n; ++i) a.ptr[i + 1] += a.ptr[i];
  std::vector<int32_t> col(E);
  std::vector<int64_t> fill(a.ptr.begin(), a.ptr.end() - 1);
  for (int64_t e = 0; e < E; ++e) col[fill[g[e]]++] = (int32_t)g[E + e];
  // sort + unique every row, then compact
  std::vector<int64_t> cnt(n, 0);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < n; ++i) {
    int32_t* b = col.data() + a.ptr[i];
    int32_t* e = col.data() + a.ptr[i + 1];
    std::sort(b, e);
    cnt[i] = std::unique(b, e) - b;
  }
  std::vector<int64_t> nptr(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) nptr[i + 1] = nptr[i] + cnt[i];
  a.col.resize(nptr[n]);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < n; ++i) memcpy(a.col.data() + nptr[i], col.data() + a.ptr[i], cnt[i] * sizeof(int32_t));
  a.ptr.swap(nptr);
  return a;
}
}  // namespace

// Weakly connected components, labelled in order of their smallest node id (graph_wrapper.py:107-134 visits nodes
// in ascending order and grows a cluster from the first unvisited one).  labels_out [n]; returns the count.
extern "C" int bsms_components_host(const int64_t* flat_edge, int64_t n_edges, int64_t n_nodes, int64_t* labels_out,
                                    int64_t* n_comp_out) {
  if (!flat_edge && n_edges > 0) return BSMS_EINVAL;
  if (!labels_out || !n_comp_out || n_nodes < 1) return BSMS_EINVAL;
  const int64_t n = n_nodes, E = n_edges;
  for (int64_t e = 0; e < 2 * E; ++e)
    if (flat_edge[e] < 0 || flat_edge[e] >= n) return BSMS_EINDEX;
  // union-find over the undirected pattern, then relabel by smallest member
  std::vector<int64_t> parent(n);
  for (int64_t i = 0; i < n; ++i) parent[i] = i;
  auto find = [&](int64_t x) {
    while (parent[x] != x) {
      parent[x] = parent[parent[x]];
      x = parent[x];
    }
    return x;
  };
  for (int64_t e = 0; e < E; ++e) {
    int64_t a = find(flat_edge[e]), b = find(flat_edge[E + e]);
    if (a != b) {
      if (a < b) parent[b] = a; else parent[a] = b;  // the root is the smallest id of the component
    }
  }
  std::vector<int64_t> lab_of_root(n, -1);
  int64_t nc = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r = find(i);
    if (lab_of_root[r] < 0) lab_of_root[r] = nc++;  // roots are met in ascending order of their (smallest) id
    labels_out[i] = lab_of_root[r];
  }
  *n_comp_out = nc;
  return BSMS_OK;
}

// One pooling level.  seeds [n_comp]: one node per cluster.  keep_out: capacity n_nodes, receives the kept node ids in
// ascending order; *edges_out receives a malloc'ed int64 [2, E'] (free with bsms_host_free).
extern "C" int bsms_bistride_level_host(const int64_t* flat_edge, int64_t n_edges, int64_t n_nodes, const int64_t* labels,
                                        int64_t n_comp, const int64_t* seeds, int64_t* keep_out, int64_t* n_keep_out,
                                        int64_t** edges_out, int64_t* n_edges_out) {
  if (!labels || !seeds || !keep_out || !n_keep_out || !edges_out || !n_edges_out || n_nodes < 1 || n_comp < 1) return BSMS_EINVAL;
  const int64_t n = n_nodes, E = n_edges;
  if (n >= (1ll << 31)) return BSMS_EINVAL;
  Csr a = build_csr(flat_edge, E, n);
  // ---- BFS depth from every cluster's seed (clusters are disjoint: one multi-source BFS), bsms_graph_wrapper.py:73-79
  std::vector<int32_t> dist(n, -1);
  std::vector<int32_t> frontier, next;
  for (int64_t c = 0; c < n_comp; ++c) {
    if (seeds[c] < 0 || seeds[c] >= n) return BSMS_EINDEX;
    if (dist[seeds[c]] < 0) {
      dist[seeds[c]] = 0;
      frontier.push_back((int32_t)seeds[c]);
    }
  }
  for (int32_t depth = 1; !frontier.empty(); ++depth) {
    next.clear();
    for (int32_t u : frontier)
      for (int64_t k = a.ptr[u]; k < a.ptr[u + 1]; ++k) {
        const int32_t v = a.col[k];
        if (dist[v] < 0) {
          dist[v] = depth;
          next.push_back(v);
        }
      }
    frontier.swap(next);
  }
  // ---- keep the smaller of the even / odd sets per cluster; even on ties or when there is no odd node (:80-95)
  std::vector<int64_t> n_even(n_comp, 0), n_odd(n_comp, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (dist[i] < 0) continue;
    if (dist[i] & 1) n_odd[labels[i]]++; else n_even[labels[i]]++;
  }
  std::vector<int32_t> new_id(n, -1);
  int64_t nk = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (dist[i] < 0) continue;
    const int64_t c = labels[i];
    const bool keep_even = n_even[c] <= n_odd[c] || n_odd[c] == 0;
    if (keep_even == !(dist[i] & 1)) {
      new_id[i] = (int32_t)nk;
      keep_out[nk++] = i;
    }
  }
  *n_keep_out = nk;
  // ---- (A+I)^2 pattern on kept rows / columns without the diagonal (:99-102, :129-154): two passes (count, fill),
  //      one stamp array per thread
  std::vector<int64_t> rptr(nk + 1, 0);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  std::vector<std::vector<int32_t>> stamp(nthreads, std::vector<int32_t>()), scratch(nthreads);
  auto row_neighbours = [&](int64_t r, std::vector<int32_t>& st, std::vector<int32_t>& out) {
    // kept columns reachable from kept node u in at most two steps of A + I, excluding u itself
    const int32_t u = (int32_t)keep_out[r];
    const int32_t tag = (int32_t)r;
    out.clear();
    st[u] = tag;
    auto visit = [&](int32_t w) {
      if (st[w] != tag) {
        st[w] = tag;
        if (new_id[w] >= 0) out.push_back(new_id[w]);
      }
    };
    for (int64_t k = a.ptr[u]; k < a.ptr[u + 1]; ++k) {
      const int32_t v = a.col[k];
      visit(v);
      for (int64_t k2 = a.ptr[v]; k2 < a.ptr[v + 1]; ++k2) visit(a.col[k2]);
    }
  };
#pragma omp parallel
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    stamp[t].assign(n, -1);
#pragma omp for schedule(dynamic, 256)
    for (int64_t r = 0; r < nk; ++r) {
      row_neighbours(r, stamp[t], scratch[t]);
      rptr[r + 1] = (int64_t)scratch[t].size();
    }
  }
  for (int64_t r = 0; r < nk; ++r) rptr[r + 1] += rptr[r];
  const int64_t En = rptr[nk];
  int64_t* eo = (int64_t*)malloc(sizeof(int64_t) * (size_t)std::max<int64_t>(2 * En, 1));
  if (!eo) return BSMS_EINVAL;
#pragma omp parallel
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    stamp[t].assign(n, -1);
#pragma omp for schedule(dynamic, 256)
    for (int64_t r = 0; r < nk; ++r) {
      row_neighbours(r, stamp[t], scratch[t]);
      std::sort(scratch[t].begin(), scratch[t].end());
      int64_t o = rptr[r];
      for (int32_t c : scratch[t]) {
        eo[o] = r;
        eo[En + o] = c;
        ++o;
      }
    }
  }
  *edges_out = eo;
  *n_edges_out = En;
  return BSMS_OK;
}

extern "C" void bsms_host_free(void* p) { free(p); }
