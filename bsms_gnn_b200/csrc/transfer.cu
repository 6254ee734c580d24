// Inter-level transfer operators: cal_ew, WeightedEdgeConv (both directions), the fused
// restriction (down-conv + pool) and prolongation (unpool + up-conv), pool/unpool row moves.
// Reference: src/ops/basic.py:101-201, src/ops/BSMS.py:73-100.  All HBM-bound gathers:
// one warp per output row, 128-bit loads, CSR rows => no atomics, deterministic.
#include "common.cuh"

namespace bsms {

// ---------------------------------------------------------------- cal_ew (basic.py:142-167)
__global__ void k_ew_node(const float* __restrict__ w, const int32_t* __restrict__ rowptr_s,
                          const int32_t* __restrict__ rowptr_d, const int32_t* __restrict__ src_d, int32_t N,
                          float* __restrict__ aggr_w) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  // sequential sum in the reference's edge order (the dst sort is stable), like CPU scatter_add_
  float acc = 0.f;
  for (int k = rowptr_d[j]; k < rowptr_d[j + 1]; ++k) {
    int i = src_d[k];
    float deg = (float)(rowptr_s[i + 1] - rowptr_s[i]);
    acc += w[i] / deg;
  }
  aggr_w[j] = acc + 1e-12f;
}
__global__ void k_ew_edge_d(const float* __restrict__ w, const float* __restrict__ aggr_w,
                            const int32_t* __restrict__ rowptr_s, const int32_t* __restrict__ src_d,
                            const int32_t* __restrict__ dst_d, const int32_t* __restrict__ perm_d, int32_t E,
                            float* __restrict__ ew_d, float* __restrict__ ew_orig) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  int i = src_d[k];
  float deg = (float)(rowptr_s[i + 1] - rowptr_s[i]);
  float v = (w[i] / deg) / aggr_w[dst_d[k]];
  ew_d[k] = v;
  if (ew_orig) ew_orig[perm_d[k]] = v;
}
__global__ void k_ew_to_s(const float* __restrict__ ew_d, const int32_t* __restrict__ s2d, int32_t E,
                          float* __restrict__ ew_s) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < E) ew_s[k] = ew_d[s2d[k]];
}
__global__ void k_ew_from_orig(const float* __restrict__ ew_orig, const int32_t* __restrict__ perm_d, int32_t E,
                               float* __restrict__ ew_d) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < E) ew_d[k] = ew_orig[perm_d[k]];
}

// ---------------------------------------------------------------- weighted CSR gather-reduce
// out[b, r, :] = sum_{k in row(node(r))} ew[k] * x[b, col(k), :]
//   node(r) = rowmap ? rowmap[r] : r         (rowmap = pooled ids: restriction forms kept rows only)
//   col(k)  = nbrmap ? nbrmap[nbr[k]] : nbr[k], skipped when < 0   (nbrmap = inverse ids: prolongation
//             reads the coarse tensor directly instead of a zero-filled fine one)
// C == 128: one warp per output row, lane owns 4 channels (one 512 B row = one coalesced request).
__global__ void __launch_bounds__(256)
k_conv128(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr, const float* __restrict__ ew,
          const int32_t* __restrict__ rowmap, const int32_t* __restrict__ nbrmap, const float* __restrict__ x,
          float* __restrict__ out, int32_t B, int32_t n_out, int32_t n_in) {
  const int lane = threadIdx.x & 31;
  long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (long long)B * n_out) return;
  int b = (int)(gw / n_out), r = (int)(gw - (long long)b * n_out);
  int node = rowmap ? rowmap[r] : r;
  if (node < 0) {  // output row without a source row (partitioned runs: coarse ghost whose fine node is remote)
    st4(out + ((size_t)b * n_out + r) * 128 + lane * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  int k0 = rowptr[node], k1 = rowptr[node + 1];
  const float* xb = x + (size_t)b * n_in * 128 + lane * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = k0; base < k1; base += 32) {
    int cnt = min(32, k1 - base);
    int my_j = -1;
    float my_w = 0.f;
    if (lane < cnt) {
      my_j = nbr[base + lane];
      if (nbrmap) my_j = nbrmap[my_j];
      my_w = ew[base + lane];
    }
    int t = 0;
    for (; t + 4 <= cnt; t += 4) {  // 4 independent 512 B row requests in flight per warp
      int j0 = __shfl_sync(0xffffffffu, my_j, t), j1 = __shfl_sync(0xffffffffu, my_j, t + 1);
      int j2 = __shfl_sync(0xffffffffu, my_j, t + 2), j3 = __shfl_sync(0xffffffffu, my_j, t + 3);
      float w0 = __shfl_sync(0xffffffffu, my_w, t), w1 = __shfl_sync(0xffffffffu, my_w, t + 1);
      float w2 = __shfl_sync(0xffffffffu, my_w, t + 2), w3 = __shfl_sync(0xffffffffu, my_w, t + 3);
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 v0 = j0 >= 0 ? ld4(xb + (size_t)j0 * 128) : z;
      float4 v1 = j1 >= 0 ? ld4(xb + (size_t)j1 * 128) : z;
      float4 v2 = j2 >= 0 ? ld4(xb + (size_t)j2 * 128) : z;
      float4 v3 = j3 >= 0 ? ld4(xb + (size_t)j3 * 128) : z;
      acc.x += w0 * v0.x; acc.y += w0 * v0.y; acc.z += w0 * v0.z; acc.w += w0 * v0.w;
      acc.x += w1 * v1.x; acc.y += w1 * v1.y; acc.z += w1 * v1.z; acc.w += w1 * v1.w;
      acc.x += w2 * v2.x; acc.y += w2 * v2.y; acc.z += w2 * v2.z; acc.w += w2 * v2.w;
      acc.x += w3 * v3.x; acc.y += w3 * v3.y; acc.z += w3 * v3.z; acc.w += w3 * v3.w;
    }
    for (; t < cnt; ++t) {
      int j = __shfl_sync(0xffffffffu, my_j, t);
      float wj = __shfl_sync(0xffffffffu, my_w, t);
      if (j >= 0) {
        float4 v = ld4(xb + (size_t)j * 128);
        acc.x += wj * v.x; acc.y += wj * v.y; acc.z += wj * v.z; acc.w += wj * v.w;
      }
    }
  }
  st4(out + ((size_t)b * n_out + r) * 128 + lane * 4, acc);
}

// any C (positions: C = pos_dim): one thread per output element
__global__ void k_conv_any(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                           const float* __restrict__ ew, const int32_t* __restrict__ rowmap,
                           const int32_t* __restrict__ nbrmap, const float* __restrict__ x, float* __restrict__ out,
                           int32_t B, int32_t n_out, int32_t n_in, int32_t C) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * n_out * C) return;
  int c = (int)(t % C);
  long long br = t / C;
  int b = (int)(br / n_out), r = (int)(br - (long long)b * n_out);
  int node = rowmap ? rowmap[r] : r;
  const float* xb = x + (size_t)b * n_in * C + c;
  float acc = 0.f;
  const int kbeg = node < 0 ? 0 : rowptr[node], kend = node < 0 ? 0 : rowptr[node + 1];
  for (int k = kbeg; k < kend; ++k) {
    int j = nbr[k];
    if (nbrmap) j = nbrmap[j];
    if (j >= 0) acc += ew[k] * xb[(size_t)j * C];
  }
  out[t] = acc;
}

// rows: out[b,k,:] = x[b,ids[k],:]
__global__ void k_gather_rows(const float* __restrict__ x, const int32_t* __restrict__ ids, int32_t n_keep,
                              int32_t n_rows, float* __restrict__ out, int32_t B, int32_t C) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * n_keep * C;
  if (t >= total) return;
  int c = (int)(t % C);
  long long bk = t / C;
  int b = (int)(bk / n_keep), k = (int)(bk - (long long)b * n_keep);
  out[t] = x[((size_t)b * n_rows + ids[k]) * C + c];
}
__global__ void k_scatter_rows(const float* __restrict__ h, const int32_t* __restrict__ ids, int32_t n_keep,
                               int32_t n_rows, float* __restrict__ out, int32_t B, int32_t C) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * n_keep * C;
  if (t >= total) return;
  int c = (int)(t % C);
  long long bk = t / C;
  int b = (int)(bk / n_keep), k = (int)(bk - (long long)b * n_keep);
  out[((size_t)b * n_rows + ids[k]) * C + c] = h[t];
}

static int launch_conv(const int32_t* rowptr, const int32_t* nbr, const float* ew, const int32_t* rowmap,
                       const int32_t* nbrmap, const float* x, float* out, int B, int n_out, int n_in, int C,
                       cudaStream_t st) {
  long long rows = (long long)B * n_out;
  if (rows == 0) return BSMS_OK;
  ProfScope ps_(PK_TRANSFER, st);
  if (C == 128) {
    long long threads = rows * 32;
    k_conv128<<<ceil_div(threads, 256), 256, 0, st>>>(rowptr, nbr, ew, rowmap, nbrmap, x, out, B, n_out, n_in);
  } else {
    long long threads = rows * C;
    k_conv_any<<<ceil_div(threads, 256), 256, 0, st>>>(rowptr, nbr, ew, rowmap, nbrmap, x, out, B, n_out, n_in, C);
  }
  BSMS_LAUNCHED();
  return BSMS_OK;
}
}  // namespace bsms

using namespace bsms;

extern "C" int bsms_cal_ew(const bsms_level_plan* p, const float* w, float* ew_orig, float* ew_d, float* ew_s,
                           float* aggr_w, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CHECK_ARG(p && w && ew_d && ew_s && aggr_w, "bsms_cal_ew: null argument");
  int N = p->n_nodes, E = p->n_edges;
  k_ew_node<<<ceil_div(N, 128), 128, 0, st>>>(w, p->rowptr_s, p->rowptr_d, p->src_d, N, aggr_w);
  BSMS_LAUNCHED();
  if (E > 0) {
    k_ew_edge_d<<<ceil_div(E, 256), 256, 0, st>>>(w, aggr_w, p->rowptr_s, p->src_d, p->dst_d, p->perm_d, E, ew_d,
                                                   ew_orig);
    BSMS_LAUNCHED();
    k_ew_to_s<<<ceil_div(E, 256), 256, 0, st>>>(ew_d, p->s2d, E, ew_s);
    BSMS_LAUNCHED();
  }
  return BSMS_OK;
}

extern "C" int bsms_permute_ew(const bsms_level_plan* p, const float* ew_orig, float* ew_d, float* ew_s,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CHECK_ARG(p != nullptr, "bsms_permute_ew: null plan");
  int E = p->n_edges;
  if (E == 0) return BSMS_OK;
  BSMS_CHECK_ARG(ew_orig && ew_d && ew_s, "bsms_permute_ew: null argument");
  k_ew_from_orig<<<ceil_div(E, 256), 256, 0, st>>>(ew_orig, p->perm_d, E, ew_d);
  BSMS_LAUNCHED();
  k_ew_to_s<<<ceil_div(E, 256), 256, 0, st>>>(ew_d, p->s2d, E, ew_s);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

extern "C" int bsms_edge_conv(const bsms_level_plan* p, const float* ew, const float* x, float* out, int32_t B,
                              int32_t C, int32_t up, void* stream) {
  BSMS_CHECK_ARG(p && x && out && (ew || p->n_edges == 0), "bsms_edge_conv: null argument");
  BSMS_CHECK_ARG(B >= 1 && C >= 1, "bsms_edge_conv: bad B/C");
  if (up)
    return launch_conv(p->rowptr_s, p->dst_s, ew, nullptr, nullptr, x, out, B, p->n_nodes, p->n_nodes, C,
                       (cudaStream_t)stream);
  return launch_conv(p->rowptr_d, p->src_d, ew, nullptr, nullptr, x, out, B, p->n_nodes, p->n_nodes, C,
                     (cudaStream_t)stream);
}

extern "C" int bsms_conv_down_pool(const bsms_level_plan* p, const float* ew_d, const int32_t* ids, int32_t n_keep,
                                   const float* x, float* out, int32_t B, int32_t C, void* stream) {
  BSMS_CHECK_ARG(p && ids && x && out && (ew_d || p->n_edges == 0), "bsms_conv_down_pool: null argument");
  return launch_conv(p->rowptr_d, p->src_d, ew_d, ids, nullptr, x, out, B, n_keep, p->n_nodes, C,
                     (cudaStream_t)stream);
}

extern "C" int bsms_unpool_conv_up(const bsms_level_plan* p, const float* ew_s, const int32_t* inv, int32_t n_keep,
                                   const float* hc, float* out, int32_t B, int32_t C, void* stream) {
  BSMS_CHECK_ARG(p && inv && hc && out && (ew_s || p->n_edges == 0), "bsms_unpool_conv_up: null argument");
  return launch_conv(p->rowptr_s, p->dst_s, ew_s, nullptr, inv, hc, out, B, p->n_nodes, n_keep, C,
                     (cudaStream_t)stream);
}

extern "C" int bsms_gather_rows(const float* x, const int32_t* ids, int32_t n_keep, int32_t n_rows, float* out,
                                int32_t B, int32_t C, void* stream) {
  BSMS_CHECK_ARG(x && ids && out, "bsms_gather_rows: null argument");
  long long total = (long long)B * n_keep * C;
  if (total == 0) return BSMS_OK;
  k_gather_rows<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(x, ids, n_keep, n_rows, out, B, C);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

extern "C" int bsms_unpool_rows(const float* h, const int32_t* ids, int32_t n_keep, int32_t n_rows, float* out,
                                int32_t B, int32_t C, void* stream) {
  BSMS_CHECK_ARG(h && ids && out, "bsms_unpool_rows: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CUDA(cudaMemsetAsync(out, 0, (size_t)B * n_rows * C * sizeof(float), st));
  long long total = (long long)B * n_keep * C;
  if (total == 0) return BSMS_OK;
  k_scatter_rows<<<ceil_div(total, 256), 256, 0, st>>>(h, ids, n_keep, n_rows, out, B, C);
  BSMS_LAUNCHED();
  return BSMS_OK;
}
