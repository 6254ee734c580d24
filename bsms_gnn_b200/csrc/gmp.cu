// GMP block (src/ops/basic.py:26-98): forward and recompute-backward orchestration plus the
// non-GEMM kernels (edge gather/fiber/combine, LayerNorm + CSR segment sum, LayerNorm backward,
// segment sums of edge gradients).  BSMS_MODE_FP32 runs the dense layers on the fp32 FFMA GEMMs in
// gemm_fp32.cuh; the tensor-core modes replace the MLP chains with the fused tcgen05 kernels in
// umma_chain.cu.
//
// Algebra used everywhere: the first edge Linear splits by column block,
//   W1 [fiber, x_i, x_j] = W1f fiber + W1s x_i + W1d x_j,
// so the two latent blocks are projected ONCE PER NODE (Ps = x W1s^T, Pd = x W1d^T) and the
// per-edge work of layer 0 is a gather-add of two projected rows plus a (P+1)-term fiber FMA.
// That removes 2*128*128*2 of the 164 608 FLOP/edge and turns the layer-0 weight gradient into
// a node-level GEMM.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gemm_fp32.cuh"

namespace bsms {

constexpr int D = BSMS_LATENT;
constexpr float LN_EPS = 1e-5f;

// ------------------------------------------------------------------------------------------
// a0[e] = relu(Ps[src_e] + Pd[dst_e] + b1 + W1f fiber_e), rows in dst-sorted edge order.
// One warp per edge row, lane owns 4 channels.  PsPd: [B*N, 256] (Ps | Pd).
// ------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(256)
k_edge_combine(const float* __restrict__ PsPd, const float* __restrict__ pos, int pos_batched,
               const int32_t* __restrict__ src_d, const int32_t* __restrict__ dst_d, const float* __restrict__ W1,
               const float* __restrict__ b1, float* __restrict__ A0, int B, int N, int E) {
  const int lane = threadIdx.x & 31;
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)B * E) return;
  int b = (int)(row / E), e = (int)(row - (long long)b * E);
  int i = src_d[e], j = dst_d[e];
  const float* pb = pos + (pos_batched ? (size_t)b * N * P : 0);
  float fib[P + 1];
  float nrm = 0.f;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    fib[p] = pb[(size_t)i * P + p] - pb[(size_t)j * P + p];
    nrm += fib[p] * fib[p];
  }
  fib[P] = sqrtf(nrm);
  const int ldw = 2 * D + P + 1;
  float4 ps = ld4(PsPd + ((size_t)b * N + i) * 256 + lane * 4);
  float4 pd = ld4(PsPd + ((size_t)b * N + j) * 256 + 128 + lane * 4);
  float4 bb = b1 ? ld4(b1 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);  // null: b1 is already folded into the Pd half
  float v[4] = {ps.x + pd.x + bb.x, ps.y + pd.y + bb.y, ps.z + pd.z + bb.z, ps.w + pd.w + bb.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float* wrow = W1 + (size_t)(lane * 4 + q) * ldw;
#pragma unroll
    for (int p = 0; p <= P; ++p) v[q] += wrow[p] * fib[p];
    v[q] = fmaxf(v[q], 0.f);
  }
  st4(A0 + row * D + lane * 4, make_float4(v[0], v[1], v[2], v[3]));
}

// warp max of |v| -> one atomicMax of its float bits into `slot` (skipped when the slot already holds as much): the
// tensor-core backward of the fp32-parity mode scales every gradient operand by a power of two derived from this
__device__ __forceinline__ void publish_amax4(unsigned* slot, const float4& v) {
  float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && __float_as_uint(m) > *reinterpret_cast<volatile unsigned*>(slot))
    atomicMax(slot, __float_as_uint(m));
}

__device__ __forceinline__ void ln_stats(const float4& y, float& mean, float& rstd) {
  float s = warp_sum(y.x + y.y + y.z + y.w);
  mean = s * (1.f / D);
  float dx = y.x - mean, dy = y.y - mean, dz = y.z - mean, dw = y.w - mean;
  float v = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / D);
  rstd = 1.f / sqrtf(v + LN_EPS);
}

// aggr[b,n,:] = sum_{k in row_d(n)} LN(Y[b*E + k, :])     (basic.py:18,94) — warp per node
__global__ void __launch_bounds__(256)
k_ln_segsum(const float* __restrict__ Y, const int32_t* __restrict__ rowptr_d, float* __restrict__ aggr, int B, int N,
            int E) {
  const int lane = threadIdx.x & 31;
  long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (long long)B * N) return;
  int b = (int)(gw / N), n = (int)(gw - (long long)b * N);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* yb = Y + (size_t)b * E * D + lane * 4;
  for (int k = rowptr_d[n]; k < rowptr_d[n + 1]; ++k) {
    float4 y = ld4(yb + (size_t)k * D);
    float mean, rstd;
    ln_stats(y, mean, rstd);
    acc.x += (y.x - mean) * rstd; acc.y += (y.y - mean) * rstd;
    acc.z += (y.z - mean) * rstd; acc.w += (y.w - mean) * rstd;
  }
  st4(aggr + gw * D + lane * 4, acc);
}

// out = LN(Yn) + x (+ skip)      (basic.py:98, BSMS.py:102) — warp per node row
__global__ void __launch_bounds__(256)
k_ln_residual(const float* __restrict__ Yn, const float* __restrict__ x, const float* __restrict__ skip,
              float* __restrict__ out, long long rows) {
  const int lane = threadIdx.x & 31;
  long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  float4 y = ld4(Yn + r * D + lane * 4);
  float mean, rstd;
  ln_stats(y, mean, rstd);
  float4 xv = x ? ld4(x + r * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);  // x == nullptr: plain LayerNorm
  float4 o = make_float4((y.x - mean) * rstd + xv.x, (y.y - mean) * rstd + xv.y, (y.z - mean) * rstd + xv.z,
                         (y.w - mean) * rstd + xv.w);
  if (skip) {
    float4 s = ld4(skip + r * D + lane * 4);
    o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
  }
  st4(out + r * D + lane * 4, o);
}

// LayerNorm backward for rows of Y: gY = rstd * (g - mean(g) - yhat * mean(g * yhat)), where the
// upstream row is g[b*N + idx[e]] for edge rows (gather of the aggregated gradient) or g[row].
__global__ void __launch_bounds__(256)
k_ln_bwd(const float* __restrict__ Y, const float* __restrict__ g, int ldg, const int32_t* __restrict__ idx,
         int rows_per_b, int g_rows_per_b, float* __restrict__ gY, long long rows, unsigned* __restrict__ amax = nullptr) {
  const int lane = threadIdx.x & 31;
  long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  long long gr = r;
  if (idx) {
    long long b = r / rows_per_b;
    gr = b * g_rows_per_b + idx[r - b * rows_per_b];
  }
  float4 y = ld4(Y + r * D + lane * 4);
  float mean, rstd;
  ln_stats(y, mean, rstd);
  float4 gv = ld4(g + gr * ldg + lane * 4);
  float4 yh = make_float4((y.x - mean) * rstd, (y.y - mean) * rstd, (y.z - mean) * rstd, (y.w - mean) * rstd);
  float c1 = warp_sum(gv.x + gv.y + gv.z + gv.w) * (1.f / D);
  float c2 = warp_sum(gv.x * yh.x + gv.y * yh.y + gv.z * yh.z + gv.w * yh.w) * (1.f / D);
  const float4 o = make_float4(rstd * (gv.x - c1 - yh.x * c2), rstd * (gv.y - c1 - yh.y * c2),
                               rstd * (gv.z - c1 - yh.z * c2), rstd * (gv.w - c1 - yh.w * c2));
  st4(gY + r * D + lane * 4, o);
  if (amax) publish_amax4(amax, o);
}

// gPsPd[b,n,0:128]  = sum_{k in row_s(n)} gU0[b*E + s2d[k]]   (gradient reaching Ps through x_i gathers)
// gPsPd[b,n,128:256]= sum_{k in row_d(n)} gU0[b*E + k]        (through x_j gathers) — warp per node
__global__ void __launch_bounds__(256)
k_edge_grad_segsum(const float* __restrict__ gU0, const int32_t* __restrict__ rowptr_d,
                   const int32_t* __restrict__ rowptr_s, const int32_t* __restrict__ s2d, float* __restrict__ gPsPd,
                   int B, int N, int E, unsigned* __restrict__ amax2 = nullptr) {
  const int lane = threadIdx.x & 31;
  long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (long long)B * N) return;
  int b = (int)(gw / N), n = (int)(gw - (long long)b * N);
  const float* gb = gU0 + (size_t)b * E * D + lane * 4;
  float4 as = make_float4(0.f, 0.f, 0.f, 0.f), ad = as;
  for (int k = rowptr_s[n]; k < rowptr_s[n + 1]; ++k) {
    float4 v = ld4(gb + (size_t)s2d[k] * D);
    as.x += v.x; as.y += v.y; as.z += v.z; as.w += v.w;
  }
  for (int k = rowptr_d[n]; k < rowptr_d[n + 1]; ++k) {
    float4 v = ld4(gb + (size_t)k * D);
    ad.x += v.x; ad.y += v.y; ad.z += v.z; ad.w += v.w;
  }
  st4(gPsPd + gw * 256 + lane * 4, as);
  st4(gPsPd + gw * 256 + 128 + lane * 4, ad);
  if (amax2) {
    publish_amax4(amax2, as);
    publish_amax4(amax2 + 1, ad);
  }
}

// gW1[:, 0:P+1] += gU0^T fiber ; gb1 += colsum(gU0).  Block = 128 threads (one per channel),
// each block walks a contiguous chunk of edge rows.
template <int P>
__global__ void __launch_bounds__(128)
k_fiber_wgrad(const float* __restrict__ gU0, const float* __restrict__ pos, int pos_batched,
              const int32_t* __restrict__ src_d, const int32_t* __restrict__ dst_d, float* __restrict__ gW1,
              float* __restrict__ gb1, int B, int N, int E, int rows_per_block, int* __restrict__ order) {
  __shared__ int s_ticket;
  const int c = threadIdx.x;
  const int bid = order ? ordered_ticket(order, &s_ticket) : (int)blockIdx.x;
  long long r0 = (long long)bid * rows_per_block;
  long long r1 = min(r0 + rows_per_block, (long long)B * E);
  float acc[P + 1], accb = 0.f;
#pragma unroll
  for (int p = 0; p <= P; ++p) acc[p] = 0.f;
  for (long long r = r0; r < r1; ++r) {
    int b = (int)(r / E), e = (int)(r - (long long)b * E);
    int i = src_d[e], j = dst_d[e];
    const float* pb = pos + (pos_batched ? (size_t)b * N * P : 0);
    float fib[P + 1], nrm = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      fib[p] = pb[(size_t)i * P + p] - pb[(size_t)j * P + p];
      nrm += fib[p] * fib[p];
    }
    fib[P] = sqrtf(nrm);
    float g = gU0[r * D + c];
    accb += g;
#pragma unroll
    for (int p = 0; p <= P; ++p) acc[p] += g * fib[p];
  }
  const int ldw = 2 * D + P + 1;
  if (order) ordered_wait(order, bid);
#pragma unroll
  for (int p = 0; p <= P; ++p) atomicAdd(&gW1[(size_t)c * ldw + p], acc[p]);
  atomicAdd(&gb1[c], accb);
  if (order) ordered_done(order);
}

// g_x = g_out + gcat[:, 0:128]   (residual + node-MLP input path); later GEMM accumulates the edge path
__global__ void k_add_rows(const float* __restrict__ a, const float* __restrict__ b, int ldb, float* __restrict__ out,
                           long long rows) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * 32) return;
  long long r = t >> 5;
  int c = (int)(t & 31) * 4;
  float4 x = ld4(a + r * D + c), y = ld4(b + r * ldb + c);
  st4(out + r * D + c, make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w));
}

int launch_ln_residual(const float* Yn, const float* x, const float* skip, float* out, long long rows, cudaStream_t st) {
  if (rows == 0) return BSMS_OK;
  ProfScope ps_(PK_OTHER, st);
  k_ln_residual<<<ceil_div(rows * 32, 256), 256, 0, st>>>(Yn, x, skip, out, rows);
  BSMS_LAUNCHED();
  return BSMS_OK;
}
int launch_ln_bwd_rows(const float* Y, const float* g, int ldg, float* gY, long long rows, cudaStream_t st, unsigned* amax = nullptr) {
  if (rows == 0) return BSMS_OK;
  ProfScope ps_(PK_LN_BWD, st);
  k_ln_bwd<<<ceil_div(rows * 32, 256), 256, 0, st>>>(Y, g, ldg, nullptr, 0, 0, gY, rows, amax);
  BSMS_LAUNCHED();
  return BSMS_OK;
}
// Y[rows,128] = (relu)(X[rows,128] W^T + bias) on the exact-fp32 FFMA GEMM (simulator.cu's fp32 mode)
int fp32_linear128(const float* X, const float* W, const float* bias, int relu, float* Y, long long rows, cudaStream_t st) {
  return gemm_nt(X, D, nullptr, 0, D, 0, W, D, bias, nullptr, 0, Y, D, rows, D, relu ? GEMM_RELU : 0, st, PK_NODE_FWD_GEMM);
}
int launch_add_rows(const float* a, const float* b, int ldb, float* out, long long rows, cudaStream_t st) {
  if (rows == 0) return BSMS_OK;
  ProfScope ps_(PK_OTHER, st);
  k_add_rows<<<ceil_div(rows * 32, 256), 256, 0, st>>>(a, b, ldb, out, rows);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

template <int P>
static int edge_combine(const float* PsPd, const float* pos, int pos_batched, const bsms_level_plan* pl,
                        const float* W1, const float* b1, float* A0, int B, cudaStream_t st) {
  long long rows = (long long)B * pl->n_edges;
  if (rows == 0) return BSMS_OK;
  {
    ProfScope ps_(PK_EDGE_COMBINE, st);
    k_edge_combine<P><<<ceil_div(rows * 32, 256), 256, 0, st>>>(PsPd, pos, pos_batched, pl->src_d, pl->dst_d, W1, b1, A0,
                                                                B, pl->n_nodes, pl->n_edges);
  }
  BSMS_LAUNCHED();
  return BSMS_OK;
}

template <int P>
static int fiber_wgrad(const float* gU0, const float* pos, int pos_batched, const bsms_level_plan* pl, float* gW1,
                       float* gb1, int B, cudaStream_t st, int* order = nullptr) {
  long long rows = (long long)B * pl->n_edges;
  if (rows == 0) return BSMS_OK;
  int rpb = (int)std::max<long long>(64, (rows + 148 * 8 - 1) / (148 * 8));
  {
    ProfScope ps_(PK_WGRAD, st);
    k_fiber_wgrad<P><<<ceil_div(rows, rpb), 128, 0, st>>>(gU0, pos, pos_batched, pl->src_d, pl->dst_d, gW1, gb1, B,
                                                          pl->n_nodes, pl->n_edges, rpb, order);
  }
  BSMS_LAUNCHED();
  return BSMS_OK;
}

#define BSMS_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != BSMS_OK) return _rc; \
  } while (0)

// ---- fp32 forward; when `keep` is set every activation is kept (backward recompute)
struct Fp32Acts {
  float *PsPd, *A0, *A1, *A2, *Y, *aggr, *N1, *N2, *N3, *Yn;
};

// fused tcgen05 edge stage (edge_chain.cu)
size_t edge_chain_pack_bytes(int mode);
int edge_chain_forward(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* PsPd, const float* pos,
                       int pos_batched, int B, int P, int mode, uint8_t* wpack, float* aggr, float* dbg, int dbg_stage,
                       cudaStream_t st, bool prepacked, uint8_t* bpack, float* eout = nullptr);

int edge_chain_backward(const bsms_level_plan* pl, const bsms_gmp_weights* w, const bsms_gmp_grads* gr, const float* PsPd,
                        const float* pos, int pos_batched, int B, int P, uint8_t* wpack, const float* g_aggr, int ld_g,
                        float* gPsPd, cudaStream_t st, bool prepacked, float* g0_rows = nullptr, float* part = nullptr);
// tensor-core orchestration of a whole GMP block (gmp_tc.cu)
int gmp_forward_tc(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos, int pos_batched,
                   const float* skip, float* out, float* saved, int B, int P, int mode, void* ws, size_t ws_bytes,
                   cudaStream_t st, const uint8_t* packed);
size_t gmp_packed_bytes();
int gmp_pack_tc(const bsms_gmp_weights* w, int P, int mode, uint8_t* packed, cudaStream_t st);
int gmp_backward_tc(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos,
                    int pos_batched, const float* saved, const float* g_out, float* g_x, const bsms_gmp_grads* gr, int B,
                    int P, void* ws, size_t ws_bytes, cudaStream_t st);

// tensor-core launchers of node_gemm.cu in the two-way fp16 split arithmetic
struct PackList;
struct WgradParams;
int lin_tc2_split(const float* X0, int ldx0, int NB, const uint8_t* const* blocks, int b_mn, const float* bias, int relu,
                  const float* mask, int ldmask, const float* add0, int ldadd0, const float* add1, int ldadd1, float* Y0,
                  int ldy0, float* Y1, int ldy1, long long rows, int kind, const unsigned* a_amax_dev, unsigned* y_amax_dev,
                  cudaStream_t st);
int wgrad_tc_split3(const float* const* G, const int* ldg, const float* const* X, const int* ldx, float* const* dW, const int* ldo,
                    float* const* db, int nprob, long long rows, const unsigned* g_amax_dev, cudaStream_t st);

// ------------------------------------------------------------------------------------------------------------------
// Backward of the fp32-parity tensor-core mode (BSMS_MODE_FP16X3) ON TENSOR CORES.  Every GEMM of the backward — the
// recomputation of the three edge layers, the data gradients and the weight gradients — runs as tcgen05 MMAs over
// two-way fp16 splits of both operands (22 significant bits, three MMAs per K step; a two-way bf16 split, 16 bits, was
// measured first and missed the 5e-4 gradient bar: 1.4e-3), with fp32 accumulation and fp32 tensors in HBM.  Operands are
// scaled by powers of two so that the fp16 pieces stay normal: weights by 2^8 and activations by 2^4 as in the forward
// (whose packed weight images are reused from `saved`); every GRADIENT tensor by its own scale, derived from the
// max |value| its producing kernel records in a device slot (gradient magnitudes follow the caller's loss scaling and
// grow through LayerNorm backward and the per-node sums, so no static scale fits).  Measured against the fp64 oracle
// level by level: every gradient within 3e-6 once the few ReLU inputs that lie within the forward rounding noise of
// zero are given the other one-sided derivative (tests/test_gpu_bsgmp.py ..._levels_kink_aware).  Per-edge activations are
// recomputed into the workspace (never kept between forward and backward); the node-level tensors come from
// `saved`.  Non-GEMM kernels (gather/combine, LayerNorm backward, segment sums) are the exact-fp32 ones of the
// fp32 mode.  Reference: src/ops/basic.py:48-98 differentiated by autograd.
static int backward_x3(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos, int pos_batched,
                       const float* saved, const float* g_out, float* g_x, const bsms_gmp_grads* gr, int B, int P, void* ws,
                       size_t ws_bytes, cudaStream_t st) {
  const int N = pl->n_nodes, E = pl->n_edges;
  const long long Rn = (long long)B * N, Re = (long long)B * E;
  const int ldw1 = 2 * D + P + 1;
  Arena sv(const_cast<float*>(saved), (size_t)-1);
  // same layout as gmp_tc.cu's carve_nodes
  const float* PsPd = sv.take<float>(Rn * 256);
  const float* aggr = sv.take<float>(Rn * D);
  const float* N1 = sv.take<float>(Rn * D);
  const float* N2 = sv.take<float>(Rn * D);
  const float* N3 = sv.take<float>(Rn * D);
  const float* Yn = sv.take<float>(Rn * D);
  Arena ar(ws, ws_bytes);
  const long long re = Re > 0 ? Re : 1;
  float* A0 = ar.take<float>(re * D);
  float* A1 = ar.take<float>(re * D);
  float* A2 = ar.take<float>(re * D);
  float* Y = ar.take<float>(re * D);
  float* Ge1 = ar.take<float>(re * D);
  float* Ge2 = ar.take<float>(re * D);
  float* Gn1 = ar.take<float>(Rn * D);
  float* Gn2 = ar.take<float>(Rn * D);
  float* g_aggr = ar.take<float>(Rn * D);
  float* gPsPd = ar.take<float>(Rn * 256);
  unsigned* am = ar.take<unsigned>(16);  // max |value| (float bits) of every gradient tensor, recorded by its producer
  if (!ar.ok()) {
    set_error("bsms_gmp_backward: workspace too small for the tensor-core fp32-parity backward");
    return BSMS_EWORKSPACE;
  }
  BSMS_CUDA(cudaMemsetAsync(am, 0, 16 * sizeof(unsigned), st));
  // the forward's fp16-split weight images ([hi | lo], 64 KB per block) sit behind the node tensors in `saved`, in the
  // block order of gmp_tc.cu: W2, W3, W4, W1s, W1d, V1a, V1b, V2, V3, V4
  const uint8_t* packs = reinterpret_cast<const uint8_t*>(sv.take<uint8_t>(1));
  enum { kW2 = 0, kW3, kW4, kW1s, kW1d, kV1a, kV1b, kV2, kV3, kV4 };
  auto blk = [&](int i) { return (const uint8_t*)(packs + (size_t)i * 65536); };
  // one 128 -> 128 layer: a_slot >= 0: the A operand is the gradient tensor whose max sits in am[a_slot], -1: an
  // activation (static scale); y_slot >= 0: record max |output| in am[y_slot]
  auto lin1 = [&](const float* X, int ldx, int wi, int b_mn, const float* bias, int relu, const float* mask, const float* add,
                  float* Yo, long long rows, int kind, int a_slot, int y_slot) {
    const uint8_t* b[1] = {blk(wi)};
    return lin_tc2_split(X, ldx, 1, b, b_mn, bias, relu, mask, D, add, D, nullptr, 0, Yo, D, nullptr, 0, rows, kind,
                         a_slot >= 0 ? am + a_slot : nullptr, y_slot >= 0 ? am + y_slot : nullptr, st);
  };
  auto wg = [&](const float* G, int ldg, const float* X, int ldx, float* dW, int ldo, float* db, long long rows, int g_slot) {
    const float* Gp[1] = {G};
    const float* Xp[1] = {X};
    float* dWp[1] = {dW};
    float* dbp[1] = {db};
    return wgrad_tc_split3(Gp, &ldg, Xp, &ldx, dWp, &ldo, dbp, 1, rows, am + g_slot, st);
  };
  // ---- node MLP backward                                                   slots: 0 Gn1, 1 Gn2, 2 Gn1', 3 G at layer-0 output
  BSMS_TRY(launch_ln_bwd_rows(Yn, g_out, D, Gn1, Rn, st, am + 0));
  BSMS_TRY(wg(Gn1, D, N3, D, gr->w_node[3], D, gr->b_node[3], Rn, 0));
  BSMS_TRY(lin1(Gn1, D, kV4, 1, nullptr, 0, N3, nullptr, Gn2, Rn, PK_DGRAD, 0, 1));
  BSMS_TRY(wg(Gn2, D, N2, D, gr->w_node[2], D, gr->b_node[2], Rn, 1));
  BSMS_TRY(lin1(Gn2, D, kV3, 1, nullptr, 0, N2, nullptr, Gn1, Rn, PK_DGRAD, 1, 2));
  BSMS_TRY(wg(Gn1, D, N1, D, gr->w_node[1], D, gr->b_node[1], Rn, 2));
  BSMS_TRY(lin1(Gn1, D, kV2, 1, nullptr, 0, N1, nullptr, Gn2, Rn, PK_DGRAD, 2, 3));  // Gn2 = gradient at the first node layer's output
  BSMS_TRY(wg(Gn2, D, x, D, gr->w_node[0], 2 * D, gr->b_node[0], Rn, 3));
  BSMS_TRY(wg(Gn2, D, aggr, D, gr->w_node[0] + D, 2 * D, nullptr, Rn, 3));
  {
    // [g_x | g_aggr] = Gn2 [V1a | V1b]; g_x also takes the residual path's g_out
    const uint8_t* b[2] = {blk(kV1a), blk(kV1b)};
    BSMS_TRY(lin_tc2_split(Gn2, D, 2, b, 1, nullptr, 0, nullptr, 0, g_out, D, nullptr, 0, g_x, D, g_aggr, D, Rn, PK_DGRAD, am + 3,
                           nullptr, st));
  }
  // ---- edge MLP: recompute a0..a2, y; LayerNorm backward; three (weight gradient, data gradient) pairs
  //                                                                           slots: 5 Ge1, 6 Ge2, 7 Ge1', 9 gPs, 10 gPd
  if (Re > 0) {
    if (P == 1) BSMS_TRY(edge_combine<1>(PsPd, pos, pos_batched, pl, w->w_edge[0], nullptr, A0, B, st));
    if (P == 2) BSMS_TRY(edge_combine<2>(PsPd, pos, pos_batched, pl, w->w_edge[0], nullptr, A0, B, st));
    if (P == 3) BSMS_TRY(edge_combine<3>(PsPd, pos, pos_batched, pl, w->w_edge[0], nullptr, A0, B, st));
    BSMS_TRY(lin1(A0, D, kW2, 0, w->b_edge[1], 1, nullptr, nullptr, A1, Re, PK_EDGE_FWD_GEMM, -1, -1));
    BSMS_TRY(lin1(A1, D, kW3, 0, w->b_edge[2], 1, nullptr, nullptr, A2, Re, PK_EDGE_FWD_GEMM, -1, -1));
    BSMS_TRY(lin1(A2, D, kW4, 0, w->b_edge[3], 0, nullptr, nullptr, Y, Re, PK_EDGE_FWD_GEMM, -1, -1));
    {
      ProfScope ps_(PK_LN_BWD, st);
      k_ln_bwd<<<ceil_div(Re * 32, 256), 256, 0, st>>>(Y, g_aggr, D, pl->dst_d, E, N, Ge1, Re, am + 5);
    }
    BSMS_LAUNCHED();
    auto dgrad_e = [&](const float* G, int wi, const float* act, float* Go, int a_slot, int y_slot) {
      return lin1(G, D, wi, 1, nullptr, 0, act, nullptr, Go, Re, PK_DGRAD, a_slot, y_slot);
    };
    auto wgrad_e = [&](const float* G, const float* act, int wl, int slot) {
      return wg(G, D, act, D, gr->w_edge[wl], D, gr->b_edge[wl], Re, slot);
    };
    BSMS_TRY(wgrad_e(Ge1, A2, 3, 5));
    BSMS_TRY(dgrad_e(Ge1, kW4, A2, Ge2, 5, 6));
    BSMS_TRY(wgrad_e(Ge2, A1, 2, 6));
    BSMS_TRY(dgrad_e(Ge2, kW3, A1, Ge1, 6, 7));
    BSMS_TRY(wgrad_e(Ge1, A0, 1, 7));
    BSMS_TRY(dgrad_e(Ge1, kW2, A0, Ge2, 7, -1));  // Ge2 = gradient at the edge input a0's pre-activation
    if (P == 1) BSMS_TRY(fiber_wgrad<1>(Ge2, pos, pos_batched, pl, gr->w_edge[0], gr->b_edge[0], B, st));
    if (P == 2) BSMS_TRY(fiber_wgrad<2>(Ge2, pos, pos_batched, pl, gr->w_edge[0], gr->b_edge[0], B, st));
    if (P == 3) BSMS_TRY(fiber_wgrad<3>(Ge2, pos, pos_batched, pl, gr->w_edge[0], gr->b_edge[0], B, st));
    {
      ProfScope ps_(PK_EDGE_GRAD_SEGSUM, st);
      k_edge_grad_segsum<<<ceil_div(Rn * 32, 256), 256, 0, st>>>(Ge2, pl->rowptr_d, pl->rowptr_s, pl->s2d, gPsPd, B, N, E, am + 9);
    }
    BSMS_LAUNCHED();
    // node-level gradients of the first edge layer: gW1s += gPs^T x, gW1d += gPd^T x, g_x += gPs W1s + gPd W1d
    const float* Gp[2] = {gPsPd, gPsPd + 128};
    const int ldg[2] = {256, 256};
    const float* Xp[2] = {x, x};
    const int ldx[2] = {D, D};
    float* dWp[2] = {gr->w_edge[0] + (P + 1), gr->w_edge[0] + (P + 1 + D)};
    const int ldo[2] = {ldw1, ldw1};
    float* dbp[2] = {nullptr, nullptr};
    BSMS_TRY(wgrad_tc_split3(Gp, ldg, Xp, ldx, dWp, ldo, dbp, 2, Rn, am + 9, st));
    BSMS_TRY(lin1(gPsPd, 256, kW1s, 1, nullptr, 0, nullptr, g_x, g_x, Rn, PK_DGRAD, 9, -1));
    BSMS_TRY(lin1(gPsPd + 128, 256, kW1d, 1, nullptr, 0, nullptr, g_x, g_x, Rn, PK_DGRAD, 10, -1));
  }
  return BSMS_OK;
}

static int fp32_forward(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos,
                        int pos_batched, int B, int P, const Fp32Acts& a, cudaStream_t st, int mode = BSMS_MODE_FP32,
                        uint8_t* wpack = nullptr) {
  const int N = pl->n_nodes, E = pl->n_edges;
  const long long Rn = (long long)B * N, Re = (long long)B * E;
  const int ldw1 = 2 * D + P + 1;
  // Ps | Pd = x [W1s ; W1d]^T  : two N=128 GEMMs writing the two halves of PsPd (ld 256)
  BSMS_TRY(gemm_nt(x, D, nullptr, 0, D, 0, w->w_edge[0] + (P + 1), ldw1, nullptr, nullptr, 0, a.PsPd, 256, Rn, D, 0, st, PK_NODE_FWD_GEMM));
  BSMS_TRY(gemm_nt(x, D, nullptr, 0, D, 0, w->w_edge[0] + (P + 1 + D), ldw1, nullptr, nullptr, 0, a.PsPd + 128, 256, Rn,
                   D, 0, st, PK_NODE_FWD_GEMM));
  if (mode != BSMS_MODE_FP32) {
    // tensor-core modes: the whole edge stage is one fused kernel that reduces into aggr
    set_error("internal: tensor-core modes are orchestrated by gmp_tc.cu");
    return BSMS_EINVAL;
  } else if (Re > 0) {
    if (P == 1) BSMS_TRY(edge_combine<1>(a.PsPd, pos, pos_batched, pl, w->w_edge[0], w->b_edge[0], a.A0, B, st));
    if (P == 2) BSMS_TRY(edge_combine<2>(a.PsPd, pos, pos_batched, pl, w->w_edge[0], w->b_edge[0], a.A0, B, st));
    if (P == 3) BSMS_TRY(edge_combine<3>(a.PsPd, pos, pos_batched, pl, w->w_edge[0], w->b_edge[0], a.A0, B, st));
    BSMS_TRY(gemm_nt(a.A0, D, nullptr, 0, D, 0, w->w_edge[1], D, w->b_edge[1], nullptr, 0, a.A1, D, Re, D, GEMM_RELU, st, PK_EDGE_FWD_GEMM));
    BSMS_TRY(gemm_nt(a.A1, D, nullptr, 0, D, 0, w->w_edge[2], D, w->b_edge[2], nullptr, 0, a.A2, D, Re, D, GEMM_RELU, st, PK_EDGE_FWD_GEMM));
    BSMS_TRY(gemm_nt(a.A2, D, nullptr, 0, D, 0, w->w_edge[3], D, w->b_edge[3], nullptr, 0, a.Y, D, Re, D, 0, st, PK_EDGE_FWD_GEMM));
  }
  if (mode == BSMS_MODE_FP32) {
    ProfScope ps_(PK_LN_SEGSUM, st);
    k_ln_segsum<<<ceil_div(Rn * 32, 256), 256, 0, st>>>(a.Y, pl->rowptr_d, a.aggr, B, N, E);
    BSMS_LAUNCHED();
  }
  BSMS_TRY(gemm_nt(x, D, a.aggr, D, D, D, w->w_node[0], 2 * D, w->b_node[0], nullptr, 0, a.N1, D, Rn, D, GEMM_RELU, st, PK_NODE_FWD_GEMM));
  BSMS_TRY(gemm_nt(a.N1, D, nullptr, 0, D, 0, w->w_node[1], D, w->b_node[1], nullptr, 0, a.N2, D, Rn, D, GEMM_RELU, st, PK_NODE_FWD_GEMM));
  BSMS_TRY(gemm_nt(a.N2, D, nullptr, 0, D, 0, w->w_node[2], D, w->b_node[2], nullptr, 0, a.N3, D, Rn, D, GEMM_RELU, st, PK_NODE_FWD_GEMM));
  BSMS_TRY(gemm_nt(a.N3, D, nullptr, 0, D, 0, w->w_node[3], D, w->b_node[3], nullptr, 0, a.Yn, D, Rn, D, 0, st, PK_NODE_FWD_GEMM));
  return BSMS_OK;
}

// Node-level tensors (PsPd, aggr, node-MLP activations) live in `saved` when the caller provides it
// (forward keeps them for backward: 3.5 KB per node row, ~9 % of the edge rows), else in the arena.
static Fp32Acts carve(Arena& ar, long long Rn, long long Re, bool keep, bool edge_bufs, float* saved) {
  Fp32Acts a;
  long long re = Re > 0 ? Re : 1;
  Arena sv(saved, saved ? (size_t)-1 : 0);
  Arena& nd = saved ? sv : ar;
  a.PsPd = nd.take<float>(Rn * 256);
  a.aggr = nd.take<float>(Rn * D);
  a.N1 = nd.take<float>(Rn * D);
  if (keep || saved) {
    a.N2 = nd.take<float>(Rn * D);
    a.N3 = nd.take<float>(Rn * D);
    a.Yn = nd.take<float>(Rn * D);
  } else {
    a.N2 = a.N3 = a.Yn = a.N1;  // in place: a CTA owns whole rows (BN = N = 128)
  }
  if (!edge_bufs) {
    a.A0 = a.A1 = a.A2 = a.Y = nullptr;  // the fused tcgen05 kernels keep every per-edge tensor on chip
  } else {
    a.A0 = ar.take<float>(re * D);
    if (keep) {
      a.A1 = ar.take<float>(re * D);
      a.A2 = ar.take<float>(re * D);
      a.Y = ar.take<float>(re * D);
    } else {
      a.A1 = a.A2 = a.Y = a.A0;
    }
  }
  return a;
}
int launch_edge_grad_segsum(const float* g0_rows, const bsms_level_plan* pl, float* gPsPd, int B, cudaStream_t st) {
  const long long Rn = (long long)B * pl->n_nodes;
  if (Rn == 0) return BSMS_OK;
  ProfScope ps_(PK_EDGE_GRAD_SEGSUM, st);
  k_edge_grad_segsum<<<ceil_div(Rn * 32, 256), 256, 0, st>>>(g0_rows, pl->rowptr_d, pl->rowptr_s, pl->s2d, gPsPd, B, pl->n_nodes,
                                                              pl->n_edges);
  BSMS_LAUNCHED();
  return BSMS_OK;
}
}  // namespace bsms

using namespace bsms;

// Deterministic option: bitwise run-to-run reproducible forward and backward.
//  * BSMS_MODE_FP32: its segment sums walk the CSR rows in order without atomics; with the switch on, its two
//    split-over-rows weight gradient kernels commit their partial sums in ticket order (gemm_fp32.cuh ordered_*).
//  * BSMS_MODE_BF16: the fused edge kernels write their per-edge-row results as rows and order-fixed CSR segment sums
//    replace the red.add reductions; every per-CTA flush becomes a partial-sum block + one ordered reduction
//    (chain.cuh "deterministic option", gmp_tc.cu).
//  * BSMS_MODE_FP16X3 is refused while the switch is on (the host routes it to BSMS_MODE_FP32, the same 1e-5 grade).
static int g_deterministic = 0;
namespace bsms {
int det_enabled() { return g_deterministic; }
size_t det_part_bytes();  // gmp_tc.cu: one partial-sum block per CTA of the widest flush
}  // namespace bsms
extern "C" void bsms_set_deterministic(int32_t on) { g_deterministic = on ? 1 : 0; }
extern "C" int32_t bsms_get_deterministic(void) { return g_deterministic; }

extern "C" size_t bsms_gmp_saved_bytes(int32_t B, int32_t N) {
  size_t Rn = (size_t)B * N;
  auto f = [](size_t n) { return align_up(n * sizeof(float), 256); };
  // node-level tensors + the packed 16-bit weight images of this GMP (the bf16 backward reuses them: 1 MB)
  return f(Rn * 256) + 5 * f(Rn * D) + 256 + (1u << 20);
}

extern "C" size_t bsms_gmp_workspace_bytes(int32_t B, int32_t N, int32_t E, int32_t mode, int32_t backward) {
  size_t Rn = (size_t)B * N, Re = (size_t)B * (E > 0 ? E : 1);
  auto f = [](size_t n) { return align_up(n * sizeof(float), 256); };
  const size_t node_bufs = f(Rn * 256) + 5 * f(Rn * D);  // PsPd, aggr, N1..N3, Yn
  const size_t scratch = 2u << 20;                        // packed weight / bias blocks
  if (mode == BSMS_MODE_BF16 || (mode == BSMS_MODE_FP16X3 && !backward)) {
    // fused tensor-core path: no per-edge buffer at all — except under the deterministic option (bf16), which moves one
    // [B*E, 128] row tensor through HBM per direction and keeps one partial-sum block per CTA
    const size_t det = (g_deterministic && mode == BSMS_MODE_BF16) ? f(Re * D) + (backward ? align_up(det_part_bytes(), 256) : 0) : 0;
    if (!backward) return node_bufs + scratch + det;
    return node_bufs + 4 * f(Rn * D) + 2 * f(Rn * 256) + scratch + det;
  }
  // fp32 path (and the fp32 backward the fp16x3 mode uses): per-edge activations are materialised
  size_t fwd = node_bufs + f(Re * D) + scratch;
  if (!backward) return fwd;
  return node_bufs + 4 * f(Re * D) + 2 * f(Re * D) + 2 * f(Rn * D) + 2 * f(Rn * 256) + scratch;
}

static int check_common(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos,
                        int B, int P, int mode) {
  BSMS_CHECK_ARG(pl && w && x && pos, "bsms_gmp: null argument");
  BSMS_CHECK_ARG(B >= 1, "bsms_gmp: B must be >= 1");
  BSMS_CHECK_ARG(P >= 1 && P <= 3, "bsms_gmp: pos_dim %d unsupported (1..3)", P);
  BSMS_CHECK_ARG(mode == BSMS_MODE_FP32 || mode == BSMS_MODE_FP16X3 || mode == BSMS_MODE_BF16, "bsms_gmp: unknown mode %d",
                 mode);
  for (int l = 0; l < 4; ++l)
    BSMS_CHECK_ARG(w->w_edge[l] && w->b_edge[l] && w->w_node[l] && w->b_node[l], "bsms_gmp: null weight %d", l);
  return BSMS_OK;
}

extern "C" size_t bsms_gmp_packed_bytes(void) { return gmp_packed_bytes(); }

extern "C" int bsms_gmp_pack(const bsms_gmp_weights* w, int32_t P, int32_t mode, void* packed, void* stream) {
  BSMS_CHECK_ARG(w && packed, "bsms_gmp_pack: null argument");
  BSMS_CHECK_ARG(P >= 1 && P <= 3, "bsms_gmp_pack: pos_dim %d unsupported (1..3)", P);
  BSMS_CHECK_ARG(mode == BSMS_MODE_FP16X3 || mode == BSMS_MODE_BF16, "bsms_gmp_pack: only the tensor-core modes pack weights");
  return gmp_pack_tc(w, P, mode, (uint8_t*)packed, (cudaStream_t)stream);
}

static int gmp_forward_impl(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos,
                            int32_t pos_batched, const float* skip, float* out, float* saved, int32_t B, int32_t P,
                            int32_t mode, void* ws, size_t ws_bytes, void* stream, const void* packed);

extern "C" int bsms_gmp_forward(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos,
                                int32_t pos_batched, const float* skip, float* out, float* saved, int32_t B, int32_t P,
                                int32_t mode, void* ws, size_t ws_bytes, void* stream) {
  return gmp_forward_impl(pl, w, x, pos, pos_batched, skip, out, saved, B, P, mode, ws, ws_bytes, stream, nullptr);
}

extern "C" int bsms_gmp_forward_packed(const bsms_level_plan* pl, const bsms_gmp_weights* w, const void* packed, const float* x,
                                       const float* pos, int32_t pos_batched, const float* skip, float* out, float* saved,
                                       int32_t B, int32_t P, int32_t mode, void* ws, size_t ws_bytes, void* stream) {
  return gmp_forward_impl(pl, w, x, pos, pos_batched, skip, out, saved, B, P, mode, ws, ws_bytes, stream, packed);
}

static int gmp_forward_impl(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos,
                            int32_t pos_batched, const float* skip, float* out, float* saved, int32_t B, int32_t P,
                            int32_t mode, void* ws, size_t ws_bytes, void* stream, const void* packed) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_TRY(check_common(pl, w, x, pos, B, P, mode));
  BSMS_CHECK_ARG(out && ws, "bsms_gmp_forward: null argument");
  if (ws_bytes < bsms_gmp_workspace_bytes(B, pl->n_nodes, pl->n_edges, mode, 0)) {
    set_error("bsms_gmp_forward: workspace too small");
    return BSMS_EWORKSPACE;
  }
  BSMS_CHECK_ARG(!g_deterministic || mode != BSMS_MODE_FP16X3,
                 "bsms_gmp_forward: the deterministic option (bsms_set_deterministic) is served by BSMS_MODE_FP32 and BSMS_MODE_BF16");
  if (mode != BSMS_MODE_FP32)
    return gmp_forward_tc(pl, w, x, pos, pos_batched, skip, out, saved, B, P, mode, ws, ws_bytes, st, (const uint8_t*)packed);
  const long long Rn = (long long)B * pl->n_nodes, Re = (long long)B * pl->n_edges;
  Arena ar(ws, ws_bytes);
  Fp32Acts a = carve(ar, Rn, Re, false, mode == BSMS_MODE_FP32, saved);
  uint8_t* wpack = ar.take<uint8_t>(edge_chain_pack_bytes(BSMS_MODE_FP16X3));
  BSMS_TRY(fp32_forward(pl, w, x, pos, pos_batched, B, P, a, st, mode, wpack));
  {
    ProfScope ps_(PK_OTHER, st);
    k_ln_residual<<<ceil_div(Rn * 32, 256), 256, 0, st>>>(a.Yn, x, skip, out, Rn);
  }
  BSMS_LAUNCHED();
  return BSMS_OK;
}

extern "C" int bsms_gmp_backward(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos,
                                 int32_t pos_batched, const float* saved, const float* g_out, float* g_x,
                                 const bsms_gmp_grads* gr, int32_t B, int32_t P, int32_t mode, void* ws, size_t ws_bytes,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_TRY(check_common(pl, w, x, pos, B, P, mode));
  BSMS_CHECK_ARG(g_out && g_x && gr && ws, "bsms_gmp_backward: null argument");
  for (int l = 0; l < 4; ++l)
    BSMS_CHECK_ARG(gr->w_edge[l] && gr->b_edge[l] && gr->w_node[l] && gr->b_node[l], "bsms_gmp_backward: null grad %d", l);
  if (ws_bytes < bsms_gmp_workspace_bytes(B, pl->n_nodes, pl->n_edges, mode, 1)) {
    set_error("bsms_gmp_backward: workspace too small");
    return BSMS_EWORKSPACE;
  }
  BSMS_CHECK_ARG(!g_deterministic || mode != BSMS_MODE_FP16X3,
                 "bsms_gmp_backward: the deterministic option (bsms_set_deterministic) is served by BSMS_MODE_FP32 and BSMS_MODE_BF16");
  if (mode == BSMS_MODE_BF16)
    return gmp_backward_tc(pl, w, x, pos, pos_batched, saved, g_out, g_x, gr, B, P, ws, ws_bytes, st);
  // fp32-parity tensor-core mode with the forward's node-level intermediates at hand: every GEMM on tcgen05
  // (BSMS_X3_BWD=ffma keeps the exact-fp32 FFMA backward below, which is also what runs without `saved`)
  static const bool x3_tc = !(getenv("BSMS_X3_BWD") && strcmp(getenv("BSMS_X3_BWD"), "ffma") == 0);
  if (mode == BSMS_MODE_FP16X3 && saved && x3_tc)
    return backward_x3(pl, w, x, pos, pos_batched, saved, g_out, g_x, gr, B, P, ws, ws_bytes, st);
  const int N = pl->n_nodes, E = pl->n_edges;
  const long long Rn = (long long)B * N, Re = (long long)B * E;
  const int ldw1 = 2 * D + P + 1;
  Arena ar(ws, ws_bytes);
  const bool fused = false;  // BSMS_MODE_BF16 was dispatched to gmp_backward_tc above; fp32 / fp16x3 backward run here
  // with `saved` from forward the fused path recomputes nothing at node level
  const bool have_saved = fused && saved != nullptr;
  Fp32Acts a = carve(ar, Rn, Re, true, !fused, have_saved ? const_cast<float*>(saved) : nullptr);
  float* Ge1 = fused ? nullptr : ar.take<float>((Re > 0 ? Re : 1) * D);
  float* Ge2 = fused ? nullptr : ar.take<float>((Re > 0 ? Re : 1) * D);
  uint8_t* wpack = ar.take<uint8_t>(edge_chain_pack_bytes(BSMS_MODE_FP16X3));
  float* Gn1 = ar.take<float>(Rn * D);
  float* Gn2 = ar.take<float>(Rn * D);
  float* gcat = ar.take<float>(Rn * 256);
  float* gPsPd = ar.take<float>(Rn * 256);
  // deterministic option: one (next ticket, committed) pair per split-over-rows weight-gradient launch
  int* order = nullptr;
  if (g_deterministic) {
    order = ar.take<int>(32);
    BSMS_CHECK_ARG(ar.ok(), "bsms_gmp_backward: workspace too small");
    BSMS_CUDA(cudaMemsetAsync(order, 0, 32 * sizeof(int), st));
  }
  auto ord = [&](int i) { return order ? order + 2 * i : nullptr; };
  // ---- recompute forward, keeping every (node-level) activation
  if (!have_saved)
    BSMS_TRY(fp32_forward(pl, w, x, pos, pos_batched, B, P, a, st, fused ? BSMS_MODE_BF16 : BSMS_MODE_FP32, wpack));
  // ---- node MLP backward
  {
    ProfScope ps_(PK_LN_BWD, st);
    k_ln_bwd<<<ceil_div(Rn * 32, 256), 256, 0, st>>>(a.Yn, g_out, D, nullptr, 0, 0, Gn1, Rn);
  }
  BSMS_LAUNCHED();
  BSMS_TRY(wgrad(Gn1, D, a.N3, D, gr->w_node[3], D, gr->b_node[3], Rn, st, ord(0)));
  BSMS_TRY(gemm_kn(Gn1, D, D, w->w_node[3], D, a.N3, D, Gn2, D, Rn, D, GEMM_MASK, st));
  BSMS_TRY(wgrad(Gn2, D, a.N2, D, gr->w_node[2], D, gr->b_node[2], Rn, st, ord(1)));
  BSMS_TRY(gemm_kn(Gn2, D, D, w->w_node[2], D, a.N2, D, Gn1, D, Rn, D, GEMM_MASK, st));
  BSMS_TRY(wgrad(Gn1, D, a.N1, D, gr->w_node[1], D, gr->b_node[1], Rn, st, ord(2)));
  BSMS_TRY(gemm_kn(Gn1, D, D, w->w_node[1], D, a.N1, D, Gn2, D, Rn, D, GEMM_MASK, st));
  // layer 0 of the node MLP: input [x | aggr]
  BSMS_TRY(wgrad(Gn2, D, x, D, gr->w_node[0], 2 * D, gr->b_node[0], Rn, st, ord(3)));
  BSMS_TRY(wgrad(Gn2, D, a.aggr, D, gr->w_node[0] + D, 2 * D, nullptr, Rn, st, ord(4)));
  BSMS_TRY(gemm_kn(Gn2, D, D, w->w_node[0], 2 * D, nullptr, 0, gcat, 256, Rn, 256, 0, st));
  {
    ProfScope ps_(PK_OTHER, st);
    k_add_rows<<<ceil_div(Rn * 32, 256), 256, 0, st>>>(g_out, gcat, 256, g_x, Rn);
  }
  BSMS_LAUNCHED();
  // ---- edge MLP backward (upstream of edge row e is g_aggr[dst_e] = gcat[:, 128:])
  if (Re > 0 && fused) {
    BSMS_CUDA(cudaMemsetAsync(gPsPd, 0, (size_t)Rn * 256 * sizeof(float), st));
    BSMS_TRY(edge_chain_backward(pl, w, gr, a.PsPd, pos, pos_batched, B, P, wpack, gcat + 128, 256, gPsPd, st, false));
  } else if (Re > 0) {
    {
      ProfScope ps_(PK_LN_BWD, st);
      k_ln_bwd<<<ceil_div(Re * 32, 256), 256, 0, st>>>(a.Y, gcat + 128, 256, pl->dst_d, E, N, Ge1, Re);
    }
    BSMS_LAUNCHED();
    BSMS_TRY(wgrad(Ge1, D, a.A2, D, gr->w_edge[3], D, gr->b_edge[3], Re, st, ord(5)));
    BSMS_TRY(gemm_kn(Ge1, D, D, w->w_edge[3], D, a.A2, D, Ge2, D, Re, D, GEMM_MASK, st));
    BSMS_TRY(wgrad(Ge2, D, a.A1, D, gr->w_edge[2], D, gr->b_edge[2], Re, st, ord(6)));
    BSMS_TRY(gemm_kn(Ge2, D, D, w->w_edge[2], D, a.A1, D, Ge1, D, Re, D, GEMM_MASK, st));
    BSMS_TRY(wgrad(Ge1, D, a.A0, D, gr->w_edge[1], D, gr->b_edge[1], Re, st, ord(7)));
    BSMS_TRY(gemm_kn(Ge1, D, D, w->w_edge[1], D, a.A0, D, Ge2, D, Re, D, GEMM_MASK, st));  // Ge2 = gU0
    if (P == 1) BSMS_TRY(fiber_wgrad<1>(Ge2, pos, pos_batched, pl, gr->w_edge[0], gr->b_edge[0], B, st, ord(8)));
    if (P == 2) BSMS_TRY(fiber_wgrad<2>(Ge2, pos, pos_batched, pl, gr->w_edge[0], gr->b_edge[0], B, st, ord(8)));
    if (P == 3) BSMS_TRY(fiber_wgrad<3>(Ge2, pos, pos_batched, pl, gr->w_edge[0], gr->b_edge[0], B, st, ord(8)));
    {
      ProfScope ps_(PK_EDGE_GRAD_SEGSUM, st);
      k_edge_grad_segsum<<<ceil_div(Rn * 32, 256), 256, 0, st>>>(Ge2, pl->rowptr_d, pl->rowptr_s, pl->s2d, gPsPd, B, N, E);
    }
    BSMS_LAUNCHED();
  }
  if (Re > 0) {
    // node-level layer-0 gradients: gW1s += gPs^T x, gW1d += gPd^T x, g_x += gPs W1s + gPd W1d
    BSMS_TRY(wgrad(gPsPd, 256, x, D, gr->w_edge[0] + (P + 1), ldw1, nullptr, Rn, st, ord(9)));
    BSMS_TRY(wgrad(gPsPd + 128, 256, x, D, gr->w_edge[0] + (P + 1 + D), ldw1, nullptr, Rn, st, ord(10)));
    BSMS_TRY(gemm_kn(gPsPd, 256, D, w->w_edge[0] + (P + 1), ldw1, nullptr, 0, g_x, D, Rn, D, GEMM_ACCUM, st));
    BSMS_TRY(gemm_kn(gPsPd + 128, 256, D, w->w_edge[0] + (P + 1 + D), ldw1, nullptr, 0, g_x, D, Rn, D, GEMM_ACCUM, st));
  }
  return BSMS_OK;
}


// Test hook: runs the pre-projection and the fused edge stage in `mode` and dumps one intermediate
// per edge row (dst-sorted order) into dbg [B*E, 128]: stage 0 = a0 (after gather/ReLU),
// 1, 2 = activations after layers 1, 2, 3 = pre-LayerNorm output of layer 3; aggr receives the result.
extern "C" int bsms_debug_edge_stage(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x,
                                     const float* pos, int32_t pos_batched, int32_t B, int32_t P, int32_t mode,
                                     int32_t stage, float* dbg, float* aggr, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_TRY(check_common(pl, w, x, pos, B, P, mode));
  BSMS_CHECK_ARG(mode != BSMS_MODE_FP32 && dbg && aggr && ws, "bsms_debug_edge_stage: bad argument");
  const long long Rn = (long long)B * pl->n_nodes;
  const int ldw1 = 2 * D + P + 1;
  Arena ar(ws, ws_bytes);
  float* PsPd = ar.take<float>(Rn * 256);
  uint8_t* wpack = ar.take<uint8_t>(edge_chain_pack_bytes(BSMS_MODE_FP16X3));
  BSMS_CHECK_ARG(ar.ok(), "bsms_debug_edge_stage: workspace too small");
  BSMS_TRY(gemm_nt(x, D, nullptr, 0, D, 0, w->w_edge[0] + (P + 1), ldw1, nullptr, nullptr, 0, PsPd, 256, Rn, D, 0, st));
  // the fused kernels expect b1 folded into the Pd half
  BSMS_TRY(gemm_nt(x, D, nullptr, 0, D, 0, w->w_edge[0] + (P + 1 + D), ldw1, w->b_edge[0], nullptr, 0, PsPd + 128, 256, Rn, D, 0, st));
  BSMS_CUDA(cudaMemsetAsync(aggr, 0, (size_t)Rn * D * sizeof(float), st));
  return edge_chain_forward(pl, w, PsPd, pos, pos_batched, B, P, mode, wpack, aggr, dbg, stage, st, false,
                            wpack + 6 * 32768);
}
