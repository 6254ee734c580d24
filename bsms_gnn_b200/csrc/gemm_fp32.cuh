// fp32 FFMA GEMMs for BSMS_MODE_FP32 (exact fp32 products, fp32 accumulate — the arithmetic of the
// reference's nn.Linear on CPU/cuBLAS-without-TF32).  128x128 CTA tile, 8x8 register tile per
// thread, BK = 16.  Three shapes cover forward, data-gradient and weight-gradient:
//   gemm_nt : Y[m,n] = act( sum_k [X|X2][m,k] * W[n,k] + bias[n] )          (Linear forward)
//   gemm_kn : Y[m,n] = ( sum_k X[m,k] * W[k,n] ) * (mask[m,n] > 0)  (+= Y)  (dgrad through ReLU)
//   wgrad   : dW[n,k] += sum_m G[m,n] * X[m,k],  db[n] += sum_m G[m,n]      (split over row chunks)
#pragma once
#include "common.cuh"

namespace bsms {

enum { GEMM_RELU = 1, GEMM_MASK = 2, GEMM_ACCUM = 4 };

constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4;

// W_KN == false: W[n*ldw + k]; true: W[k*ldw + n]
template <bool W_KN>
__global__ void __launch_bounds__(256)
k_gemm(const float* __restrict__ X, int ldx, const float* __restrict__ X2, int ldx2, int K1, int K2,
       const float* __restrict__ W, int ldw, const float* __restrict__ bias, const float* __restrict__ mask, int ldmask,
       float* __restrict__ Y, int ldy, long long M, int flags) {
  __shared__ __align__(16) float As[GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * GBM;
  const int n0 = blockIdx.y * GBN;
  const int K = K1 + K2;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += GBK) {
    const float* xs = (k0 < K1) ? X : X2;
    const int lds = (k0 < K1) ? ldx : ldx2;
    const int kk = (k0 < K1) ? k0 : k0 - K1;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int r = (tid >> 2) + it * 64, kq = (tid & 3) * 4;
      long long m = m0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) v = ld4(xs + m * lds + kk + kq);
      As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    }
    if (!W_KN) {
      int n = tid >> 1, kh = (tid & 1) * 8;
      const float* wp = W + (size_t)(n0 + n) * ldw + k0 + kh;
#pragma unroll
      for (int q = 0; q < 8; ++q) Bs[kh + q][n] = wp[q];
    } else {
      int k = tid >> 4, nq = (tid & 15) * 8;
      const float* wp = W + (size_t)(k0 + k) * ldw + n0 + nq;
#pragma unroll
      for (int q = 0; q < 8; ++q) Bs[k][nq + q] = wp[q];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int n = n0 + jh * 64 + tx * 4;
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (bias) {
        float4 bb = ld4(bias + n);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      if (flags & GEMM_RELU) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      if (flags & GEMM_MASK) {
        float4 mk = ld4(mask + m * ldmask + n);
        v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f;
        v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
      }
      float* yp = Y + m * ldy + n;
      if (flags & GEMM_ACCUM) {
        float4 o = ld4(yp);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      st4(yp, v);
    }
  }
}

static inline int gemm_nt(const float* X, int ldx, const float* X2, int ldx2, int K1, int K2, const float* W, int ldw,
                          const float* bias, const float* mask, int ldmask, float* Y, int ldy, long long M, int N,
                          int flags, cudaStream_t st, int kind = PK_OTHER) {
  if (M == 0) return BSMS_OK;
  dim3 grid(ceil_div(M, GBM), N / GBN);
  ProfScope ps(kind, st);
  k_gemm<false><<<grid, 256, 0, st>>>(X, ldx, X2, ldx2, K1, K2, W, ldw, bias, mask, ldmask, Y, ldy, M, flags);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

static inline int gemm_kn(const float* X, int ldx, int K, const float* W, int ldw, const float* mask, int ldmask,
                          float* Y, int ldy, long long M, int N, int flags, cudaStream_t st, int kind = PK_DGRAD) {
  if (M == 0) return BSMS_OK;
  dim3 grid(ceil_div(M, GBM), N / GBN);
  ProfScope ps(kind, st);
  k_gemm<true><<<grid, 256, 0, st>>>(X, ldx, nullptr, 0, K, 0, W, ldw, nullptr, mask, ldmask, Y, ldy, M, flags);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

// Ordered commit of per-CTA partial sums (the deterministic option, bsms_set_deterministic): a CTA takes a ticket when
// it STARTS (so every lower ticket is already running: no dependence on the dispatch order), works on the chunk of that
// ticket, and adds its partial sums only after every lower ticket has committed — each address then receives its
// addends in one fixed order.  t[0]: next ticket, t[1]: tickets committed (both zeroed by the caller).
__device__ __forceinline__ int ordered_ticket(int* t, int* s_slot) {
  if (threadIdx.x == 0) *s_slot = atomicAdd(&t[0], 1);
  __syncthreads();
  return *s_slot;
}
__device__ __forceinline__ void ordered_wait(int* t, int ticket) {
  if (threadIdx.x == 0) {
    while (atomicAdd(&t[1], 0) != ticket) __nanosleep(64);
  }
  __syncthreads();
}
__device__ __forceinline__ void ordered_done(int* t) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&t[1], 1);
}

// dW[n*ldo + k] += sum_{m in chunk} G[m*ldg + n] * X[m*ldx + k]   (n, k in [0,128))
__global__ void __launch_bounds__(256)
k_wgrad(const float* __restrict__ G, int ldg, const float* __restrict__ X, int ldx, float* __restrict__ dW, int ldo,
        float* __restrict__ db, long long M, int rows_per_cta, int* __restrict__ order) {
  __shared__ __align__(16) float Gs[GBK][128 + GPAD];
  __shared__ __align__(16) float Xs[GBK][128 + GPAD];
  __shared__ int s_ticket;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int bid = order ? ordered_ticket(order, &s_ticket) : (int)blockIdx.x;
  long long r0 = (long long)bid * rows_per_cta;
  long long r1 = min(r0 + rows_per_cta, M);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  for (long long rb = r0; rb < r1; rb += GBK) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int r = (tid >> 5) + it * 8, c = (tid & 31) * 4;
      long long m = rb + r;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f), x = g;
      if (m < r1) {
        g = ld4(G + m * ldg + c);
        x = ld4(X + m * ldx + c);
      }
      *reinterpret_cast<float4*>(&Gs[r][c]) = g;
      *reinterpret_cast<float4*>(&Xs[r][c]) = x;
    }
    __syncthreads();
    if (db && tid < 128) {
#pragma unroll
      for (int r = 0; r < GBK; ++r) bsum += Gs[r][tid];
    }
#pragma unroll
    for (int r = 0; r < GBK; ++r) {
      float4 a0 = *reinterpret_cast<const float4*>(&Gs[r][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&Gs[r][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Xs[r][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Xs[r][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (order) ordered_wait(order, bid);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int n = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      atomicAdd(&dW[(size_t)n * ldo + k], acc[i][j]);
    }
  }
  if (db && tid < 128) atomicAdd(&db[tid], bsum);
  if (order) ordered_done(order);
}

// order: nullptr, or two zeroed ints for the ordered (deterministic) commit
static inline int wgrad(const float* G, int ldg, const float* X, int ldx, float* dW, int ldo, float* db, long long M,
                        cudaStream_t st, int* order = nullptr) {
  if (M == 0) return BSMS_OK;
  long long per = (M + 148 * 4 - 1) / (148 * 4);
  per = (per + GBK - 1) / GBK * GBK;
  if (per < 256) per = 256;
  ProfScope ps(PK_WGRAD, st);
  k_wgrad<<<ceil_div(M, per), 256, 0, st>>>(G, ldg, X, ldx, dW, ldo, db, M, (int)per, order);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

}  // namespace bsms
