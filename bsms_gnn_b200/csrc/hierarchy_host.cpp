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
#include <sys/mman.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/bsms_b200.h"

namespace {
struct Tick {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  bool on = getenv("BSMS_HIER_PROF") != nullptr;
  void lap(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "  [hier] %-12s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};
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
  Tick tk;
  Csr a = build_csr(flat_edge, E, n);
  tk.lap("csr");
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
  tk.lap("bfs");
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
  tk.lap("keep");
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
  tk.lap("count");
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
  tk.lap("fill");
  *edges_out = eo;
  *n_edges_out = En;
  return BSMS_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// Whole hierarchy in one call.  Same result as `depth` rounds of bsms_components_host + (numpy seed choice) +
// bsms_bistride_level_host, but the graph stays an int32 CSR between levels (every level's output rows are already
// sorted and unique), clusters come from a lock-free parallel union-find, the seed choice runs here in the caller's
// floating-point type with numpy's operation order (see seeds_of below), the two-step neighbourhoods are formed in ONE
// pass from kept-only neighbour lists into per-chunk buffers, and the int64 edge lists the caller wants are written
// by all threads.  2.0 M nodes / 12 M edges / depth 6: see DESIGN.md §6.
namespace {

inline int64_t ceil_div_i64(int64_t a, int64_t b) { return (a + b - 1) / b; }
// uninitialised, 2 MB-aligned, transparent huge pages requested (fresh-page faults dominate otherwise)
inline void* big_alloc(size_t bytes) {
  const size_t two_mb = (size_t)2 << 20;
  if (bytes < two_mb) return malloc(std::max<size_t>(bytes, 1));
  void* p = nullptr;
  if (posix_memalign(&p, two_mb, (bytes + two_mb - 1) / two_mb * two_mb) != 0) return nullptr;
#ifdef MADV_HUGEPAGE
  madvise(p, (bytes + two_mb - 1) / two_mb * two_mb, MADV_HUGEPAGE);
#endif
  return p;
}

inline int32_t uf_find(int32_t* P, int32_t x) {
  for (;;) {
    int32_t p = __atomic_load_n(&P[x], __ATOMIC_RELAXED);
    if (p == x) return x;
    int32_t gp = __atomic_load_n(&P[p], __ATOMIC_RELAXED);
    if (gp != p) __atomic_compare_exchange_n(&P[x], &p, gp, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED);  // path halving
    x = p;
  }
}
// parents always point at smaller ids, so the root of a set is its smallest member and there are no cycles
inline void uf_unite(int32_t* P, int32_t a, int32_t b) {
  for (;;) {
    a = uf_find(P, a);
    b = uf_find(P, b);
    if (a == b) return;
    if (a > b) std::swap(a, b);
    int32_t expect = b;
    if (__atomic_compare_exchange_n(&P[b], &expect, a, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return;
  }
}

// Per cluster the node nearest the cluster centroid (bsms_graph_wrapper.py:107-126), in the arithmetic of the numpy
// expressions the reference evaluates — exactness of the hierarchy depends on it:
//   center = np.mean(pos_c, axis=0)      : per coordinate a SEQUENTIAL sum over the members in ascending order in the
//                                          array's own type (numpy sums along the slow axis row by row; its pairwise
//                                          summation applies to the contiguous axis only), then one division by the count
//   d = np.linalg.norm(pos_c - center, 2, axis=-1) : sqrt(x0*x0 + x1*x1 (+ x2*x2)), left to right, same type
//   seed = members[np.argmin(d)]         : first minimum
// orig[i] = level-0 id of node i (positions are read through it).  Compiled with -ffp-contract=off.
template <typename T>
void seeds_of(const T* pos, int P, const int32_t* orig, const int32_t* labels, int64_t n, int64_t nc,
              std::vector<int32_t>& seeds) {
  std::vector<T> sum((size_t)nc * P, (T)0);
  std::vector<int64_t> cnt(nc, 0);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t c = labels[i];
    const T* p = pos + (size_t)orig[i] * P;
    for (int j = 0; j < P; ++j) sum[c * P + j] += p[j];
    cnt[c]++;
  }
  for (int64_t c = 0; c < nc; ++c)
    for (int j = 0; j < P; ++j) sum[c * P + j] = sum[c * P + j] / (T)cnt[c];
  std::vector<T> best(nc, (T)0);
  seeds.assign(nc, -1);
  if (nc == 1) {
    // one cluster: the distance pass is the expensive part, run it over thread chunks (first minimum wins ties)
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    std::vector<T> tb(nt, (T)0);
    std::vector<int64_t> ti(nt, -1);
#pragma omp parallel
    {
      int t = 0;
#ifdef _OPENMP
      t = omp_get_thread_num();
#endif
      T b = (T)0;
      int64_t bi = -1;
#pragma omp for schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        const T* p = pos + (size_t)orig[i] * P;
        T d2 = (T)0;
        for (int j = 0; j < P; ++j) {
          const T x = p[j] - sum[j];
          d2 = j == 0 ? x * x : d2 + x * x;
        }
        const T d = std::sqrt(d2);
        if (bi < 0 || d < b) { b = d; bi = i; }
      }
      tb[t] = b;
      ti[t] = bi;
    }
    int64_t bi = -1;
    T b = (T)0;
    for (int t = 0; t < nt; ++t)  // static schedule: thread t owns an earlier index range than thread t + 1
      if (ti[t] >= 0 && (bi < 0 || tb[t] < b)) { b = tb[t]; bi = ti[t]; }
    seeds[0] = (int32_t)bi;
    return;
  }
  for (int64_t i = 0; i < n; ++i) {
    const int64_t c = labels[i];
    const T* p = pos + (size_t)orig[i] * P;
    T d2 = (T)0;
    for (int j = 0; j < P; ++j) {
      const T x = p[j] - sum[c * P + j];
      d2 = j == 0 ? x * x : d2 + x * x;
    }
    const T d = std::sqrt(d2);
    if (seeds[c] < 0 || d < best[c]) { best[c] = d; seeds[c] = (int32_t)i; }
  }
}

// malloc'ed, NOT zero-filled (a value-initialised std::vector would touch 128 MB per level on one thread first)
struct RawI64 {
  int64_t* p = nullptr;
  size_t n = 0;
  RawI64() = default;
  explicit RawI64(size_t count) : p((int64_t*)big_alloc(sizeof(int64_t) * std::max<size_t>(count, 1))), n(count) {}
  RawI64(RawI64&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  RawI64(const RawI64&) = delete;
  RawI64& operator=(const RawI64&) = delete;
  ~RawI64() { free(p); }
};
struct Hierarchy {
  std::vector<RawI64> edges;                // level l >= 1: [2, E_l] row-major (rows ascending, columns sorted)
  std::vector<std::vector<int64_t>> ids;    // level l >= 1: kept node ids of level l - 1, ascending
};

// CSR of the level-0 edge list with the index range check in the same pass.  `sorted` inputs (rows ascending, columns
// strictly ascending inside a row — what np.unique-based mesh generators and this builder emit) are converted without
// the sort / unique pass.  Returns BSMS_EINDEX for an out-of-range index.
int csr_level0(const int64_t* g, int64_t E, int64_t n, Csr& a) {
  bool sorted = true, ok = true;
  a.col.resize(E);
#pragma omp parallel for schedule(static) reduction(&& : sorted, ok)
  for (int64_t e = 0; e < E; ++e) {
    const int64_t r = g[e], c = g[E + e];
    ok = ok && r >= 0 && r < n && c >= 0 && c < n;
    a.col[e] = (int32_t)c;
    if (e > 0) sorted = sorted && (g[e - 1] < r || (g[e - 1] == r && g[E + e - 1] < c));
  }
  if (!ok) return BSMS_EINDEX;
  if (!sorted) {
    a = build_csr(g, E, n);
    return BSMS_OK;
  }
  a.ptr.assign(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < E; ++e) {
    const int64_t r = g[e], pr = e ? g[e - 1] : -1;
    for (int64_t q = pr + 1; q <= r; ++q) a.ptr[q] = e;  // rows (g[e-1], g[e]] start at e (sorted: each row once)
  }
  for (int64_t q = (E ? g[E - 1] + 1 : 0); q <= n; ++q) a.ptr[q] = E;
  return BSMS_OK;
}

// Grow-only scratch that survives the level loop: a level's buffers are the previous level's, so only level 0 pays
// the page faults of fresh memory (they cost more than the arithmetic here); 2 MB-aligned with MADV_HUGEPAGE.
template <typename V>
struct Buf {
  V* p = nullptr;
  size_t cap = 0;
  Buf() = default;
  Buf(const Buf&) = delete;
  Buf& operator=(const Buf&) = delete;
  ~Buf() { free(p); }
  V* ensure(size_t count) {
    if (count > cap) {
      free(p);
      p = (V*)big_alloc(count * sizeof(V));
      cap = p ? count : 0;
    }
    return p;
  }
  void swap(Buf& o) {
    std::swap(p, o.p);
    std::swap(cap, o.cap);
  }
};

template <typename T>
int build_levels(Csr a0, int64_t n, const T* pos, int P, int depth, Hierarchy& H) {
  Tick tk;
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  // current / next level graph (int32 columns), moved out of the level-0 vectors once
  Buf<int64_t> aptr, bptr, kptr;
  Buf<int32_t> acol, bcol, kcol, orig, norig, labels, ufp, new_id;
  Buf<uint8_t> par;
  int64_t E = (int64_t)a0.col.size();
  if (!aptr.ensure(n + 1) || !acol.ensure(std::max<int64_t>(E, 1)) || !orig.ensure(n)) return BSMS_EINVAL;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i <= n; ++i) aptr.p[i] = a0.ptr[i];
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < E; ++e) acol.p[e] = a0.col[e];
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) orig.p[i] = (int32_t)i;
  { Csr drop; std::swap(drop, a0); }
  std::vector<std::vector<uint64_t>> bitmaps(nthreads);
  std::vector<std::vector<int32_t>> ccol;
  std::vector<int32_t> seeds, frontier, next;
  std::vector<int64_t> n_even, n_odd;
  int64_t* emit_to = nullptr;  // the int64 [2, E] copy of the CURRENT level still to be written (levels >= 1)
  tk.lap("setup");

  // rows [e0, e1) of the current level's edge list -> emit_to
  auto emit_range = [&](int64_t e0, int64_t e1) {
    int64_t r = std::upper_bound(aptr.p, aptr.p + n + 1, e0) - aptr.p - 1;
    for (int64_t o = e0; o < e1; ++o) {
      while (o >= aptr.p[r + 1]) ++r;
      emit_to[o] = r;
      emit_to[E + o] = acol.p[o];
    }
  };
  const int64_t kEmitChunk = 1 << 16;

  for (int lvl = 0; lvl < depth; ++lvl) {
    // ---- weakly connected clusters, labelled in ascending order of their smallest node (graph_wrapper.py:107-134)
    if (!ufp.ensure(n) || !labels.ensure(n) || !par.ensure(n) || !new_id.ensure(n) || !kptr.ensure(n + 1)) return BSMS_EINVAL;
    int32_t* UF = ufp.p;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      UF[i] = (int32_t)i;
      par.p[i] = 0xFF;
      new_id.p[i] = -1;
    }
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < n; ++i)
      for (int64_t k = aptr.p[i]; k < aptr.p[i + 1]; ++k) uf_unite(UF, (int32_t)i, acol.p[k]);
    int64_t nc = 0;
    for (int64_t i = 0; i < n; ++i) {  // a root is the smallest member: it is met before every other member
      const int32_t r = uf_find(UF, (int32_t)i);
      labels.p[i] = r == i ? (int32_t)nc++ : labels.p[r];
    }
    tk.lap("components");
    seeds_of<T>(pos, P, orig.p, labels.p, n, nc, seeds);
    tk.lap("seeds");
    // ---- BFS parity from every cluster's seed (bsms_graph_wrapper.py:73-79); 0xFF = not reached.  Frontiers of a
    //      mesh are narrow (~sqrt(n)), so the BFS stays on one thread — and the other threads write the caller's int64
    //      copy of this level meanwhile
    frontier.clear();
    for (int64_t c = 0; c < nc; ++c)
      if (par.p[seeds[c]] == 0xFF) {
        par.p[seeds[c]] = 0;
        frontier.push_back(seeds[c]);
      }
    {
      int64_t emit_next = 0;
      const int64_t emit_chunks = emit_to ? ceil_div_i64(E, kEmitChunk) : 0;
#pragma omp parallel
      {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        if (t == 0) {
          for (uint8_t pd = 1; !frontier.empty(); pd ^= 1) {
            next.clear();
            const size_t nf = frontier.size();
            for (size_t f = 0; f < nf; ++f) {
              if (f + 8 < nf) __builtin_prefetch(&aptr.p[frontier[f + 8]]);
              if (f + 4 < nf) __builtin_prefetch(&acol.p[aptr.p[frontier[f + 4]]]);
              const int32_t u = frontier[f];
              for (int64_t k = aptr.p[u]; k < aptr.p[u + 1]; ++k) {
                const int32_t v = acol.p[k];
                if (par.p[v] == 0xFF) {
                  par.p[v] = pd;
                  next.push_back(v);
                }
              }
            }
            frontier.swap(next);
          }
        }
        for (;;) {
          const int64_t c = __atomic_fetch_add(&emit_next, 1, __ATOMIC_RELAXED);
          if (c >= emit_chunks) break;
          emit_range(c * kEmitChunk, std::min(E, (c + 1) * kEmitChunk));
        }
      }
      emit_to = nullptr;
    }
    tk.lap("bfs | emit");
    // ---- keep the smaller of the even / odd sets per cluster; even on ties or when there is no odd node (:80-95)
    n_even.assign(nc, 0);
    n_odd.assign(nc, 0);
    for (int64_t i = 0; i < n; ++i) {
      if (par.p[i] == 0xFF) continue;
      if (par.p[i]) n_odd[labels.p[i]]++; else n_even[labels.p[i]]++;
    }
    std::vector<int64_t> keep;
    keep.reserve(n / 2 + 16);
    for (int64_t i = 0; i < n; ++i) {
      if (par.p[i] == 0xFF) continue;
      const int64_t c = labels.p[i];
      const bool keep_even = n_even[c] <= n_odd[c] || n_odd[c] == 0;
      if (keep_even == (par.p[i] == 0)) {
        new_id.p[i] = (int32_t)keep.size();
        keep.push_back(i);
      }
    }
    const int64_t nk = (int64_t)keep.size();
    tk.lap("keep");
    // ---- kept-only neighbour lists of A + I (new ids, ascending): half the entries of A on a bi-stride level
    const int32_t* nid = new_id.p;
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < n; ++v) {
      int64_t c = nid[v] >= 0 ? 1 : 0;
      for (int64_t k = aptr.p[v]; k < aptr.p[v + 1]; ++k) c += (nid[acol.p[k]] >= 0 && acol.p[k] != v);
      kptr.p[v + 1] = c;
    }
    kptr.p[0] = 0;
    for (int64_t v = 0; v < n; ++v) kptr.p[v + 1] += kptr.p[v];
    if (!kcol.ensure(std::max<int64_t>(kptr.p[n], 1))) return BSMS_EINVAL;
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < n; ++v) {
      int64_t o = kptr.p[v];
      bool self_done = nid[v] < 0;
      for (int64_t k = aptr.p[v]; k < aptr.p[v + 1]; ++k) {
        const int32_t w = acol.p[k];
        if (w == v) continue;
        if (!self_done && w > v) {
          kcol.p[o++] = nid[v];
          self_done = true;
        }
        if (nid[w] >= 0) kcol.p[o++] = nid[w];
      }
      if (!self_done) kcol.p[o++] = nid[v];
    }
    tk.lap("kept lists");
    // ---- pattern of (A+I)^2 on kept rows / columns without the diagonal (:99-102, :129-154), one pass: row r (kept
    //      node u) = union of the kept lists of u and of every neighbour of u, minus r itself
    const int64_t nchunks = std::min<int64_t>(std::max<int64_t>(nk, 1), (int64_t)nthreads * 16);
    if ((int64_t)ccol.size() < nchunks) ccol.resize(nchunks);
    if (!bptr.ensure(nk + 1)) return BSMS_EINVAL;
    int64_t* rlen = bptr.p;
    rlen[0] = 0;
#pragma omp parallel
    {
      int t = 0;
#ifdef _OPENMP
      t = omp_get_thread_num();
#endif
      // one bit per kept node, all zero between rows.  The inner loop only ORs bits (no compare, no branch); a row is
      // then read back in ascending order from the words between the smallest and the largest id it touched — mesh
      // neighbourhoods are local in the id space, so that span is a few dozen words — and the words are cleared on
      // the way: no stamp array (32x the cache footprint), no per-row sort
      std::vector<uint64_t>& bm = bitmaps[t];
      const size_t words = (size_t)(nk + 63) / 64 + 1;
      if (bm.size() < words) bm.resize(words);
      std::fill(bm.begin(), bm.begin() + words, 0ull);
#pragma omp for schedule(dynamic, 1)
      for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t r0 = nk * c / nchunks, r1 = nk * (c + 1) / nchunks;
        std::vector<int32_t>& out = ccol[c];
        out.clear();
        for (int64_t r = r0; r < r1; ++r) {
          const int32_t u = (int32_t)keep[r];
          int32_t lo = (int32_t)r, hi = (int32_t)r;
          auto take = [&](int32_t v) {
            const int64_t k0 = kptr.p[v], k1 = kptr.p[v + 1];
            if (k0 == k1) return;
            lo = std::min(lo, kcol.p[k0]);       // the lists are ascending
            hi = std::max(hi, kcol.p[k1 - 1]);
            for (int64_t k = k0; k < k1; ++k) {
              const int32_t w = kcol.p[k];
              bm[(size_t)w >> 6] |= 1ull << (w & 63);
            }
          };
          take(u);
          for (int64_t k = aptr.p[u]; k < aptr.p[u + 1]; ++k)
            if (acol.p[k] != u) take(acol.p[k]);
          bm[(size_t)r >> 6] &= ~(1ull << (r & 63));  // no diagonal
          const size_t start = out.size();
          for (size_t wd = (size_t)lo >> 6; wd <= ((size_t)hi >> 6); ++wd) {
            uint64_t bits = bm[wd];
            if (!bits) continue;
            bm[wd] = 0;
            const int32_t base = (int32_t)(wd << 6);
            while (bits) {
              out.push_back(base + __builtin_ctzll(bits));
              bits &= bits - 1;
            }
          }
          rlen[r + 1] = (int64_t)(out.size() - start);
        }
      }
    }
    tk.lap("square");
    for (int64_t r = 0; r < nk; ++r) rlen[r + 1] += rlen[r];
    const int64_t En = rlen[nk];
    if (!bcol.ensure(std::max<int64_t>(En, 1)) || !norig.ensure(std::max<int64_t>(nk, 1))) return BSMS_EINVAL;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < nchunks; ++c) {
      const int64_t r0 = nk * c / nchunks, r1 = nk * (c + 1) / nchunks;
      if (r1 > r0) memcpy(bcol.p + bptr.p[r0], ccol[c].data(), (size_t)(bptr.p[r1] - bptr.p[r0]) * sizeof(int32_t));
    }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nk; ++r) norig.p[r] = orig.p[keep[r]];
    orig.swap(norig);
    aptr.swap(bptr);
    acol.swap(bcol);
    H.edges.emplace_back((size_t)(2 * En));
    if (!H.edges.back().p) return BSMS_EINVAL;
    emit_to = H.edges.back().p;  // written during the next level's BFS (or after the loop)
    H.ids.emplace_back(std::move(keep));
    n = nk;
    E = En;
    tk.lap("assemble");
    if (n == 0) {
      // nothing left to coarsen: the remaining levels are empty (the numpy builder yields the same)
      for (int l2 = lvl + 1; l2 < depth; ++l2) {
        H.edges.emplace_back((size_t)0);
        H.ids.emplace_back();
      }
      break;
    }
  }
  if (emit_to) {
    const int64_t emit_chunks = ceil_div_i64(E, kEmitChunk);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < emit_chunks; ++c) emit_range(c * kEmitChunk, std::min(E, (c + 1) * kEmitChunk));
    tk.lap("last emit");
  }
  return BSMS_OK;
}
}  // namespace

extern "C" int bsms_hierarchy_build_host(const int64_t* flat_edge, int64_t n_edges, int64_t n_nodes, const void* pos,
                                         int32_t pos_dim, int32_t pos_is_f64, int32_t depth, void** handle_out) {
  if ((!flat_edge && n_edges > 0) || !pos || !handle_out || n_nodes < 1 || n_nodes >= (1ll << 31) || pos_dim < 1 || depth < 0)
    return BSMS_EINVAL;
  const int64_t n = n_nodes, E = n_edges;
  Tick tk;
  Csr a;
  int rc0 = csr_level0(flat_edge, E, n, a);
  if (rc0 != BSMS_OK) return rc0;
  tk.lap("check + csr0");
  Hierarchy* H = new Hierarchy();
  int rc = pos_is_f64 ? build_levels<double>(std::move(a), n, (const double*)pos, pos_dim, depth, *H)
                      : build_levels<float>(std::move(a), n, (const float*)pos, pos_dim, depth, *H);
  if (rc != BSMS_OK) {
    delete H;
    return rc;
  }
  *handle_out = H;
  return BSMS_OK;
}

// level in [1, depth]: the level's edge list [2, E] and the ids of its nodes in level - 1; the pointers stay valid
// until bsms_hierarchy_free_host
extern "C" int bsms_hierarchy_level_host(void* handle, int32_t level, int64_t* n_nodes_out, int64_t* n_edges_out,
                                         const int64_t** edges_out, const int64_t** ids_out) {
  Hierarchy* H = (Hierarchy*)handle;
  if (!H || level < 1 || level > (int32_t)H->ids.size() || !n_nodes_out || !n_edges_out || !edges_out || !ids_out) return BSMS_EINVAL;
  *n_nodes_out = (int64_t)H->ids[level - 1].size();
  *n_edges_out = (int64_t)H->edges[level - 1].n / 2;
  *edges_out = H->edges[level - 1].p;
  *ids_out = H->ids[level - 1].data();
  return BSMS_OK;
}

extern "C" void bsms_hierarchy_free_host(void* handle) { delete (Hierarchy*)handle; }

extern "C" void bsms_host_free(void* p) { free(p); }
