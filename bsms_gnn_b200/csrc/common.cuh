// Shared host/device helpers for libbsms_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/bsms_b200.h"

namespace bsms {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define BSMS_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      bsms::set_error(__VA_ARGS__);          \
      return BSMS_EINVAL;                    \
    }                                        \
  } while (0)

#define BSMS_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      bsms::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return BSMS_ECUDA;                                                             \
    }                                                                                \
  } while (0)

// call after every kernel launch: counts it and surfaces launch-configuration errors
#define BSMS_LAUNCHED()                     \
  do {                                      \
    bsms::g_launches.fetch_add(1);          \
    BSMS_CUDA(cudaGetLastError());          \
  } while (0)

// ---- optional per-kernel CUDA-event timing (bench.py's roofline pass; off by default)
enum ProfKind {
  PK_EDGE_FWD_GEMM = 0,  // edge-MLP layers, forward (or the fused edge chain)
  PK_NODE_FWD_GEMM = 1,  // node-level GEMMs, forward (pre-projection + node MLP)
  PK_EDGE_COMBINE = 2,   // gather Ps/Pd + fiber + ReLU
  PK_LN_SEGSUM = 3,      // LayerNorm + CSR segment sum (aggregation)
  PK_DGRAD = 4,
  PK_WGRAD = 5,
  PK_LN_BWD = 6,
  PK_EDGE_GRAD_SEGSUM = 7,
  PK_TRANSFER = 8,  // restriction / prolongation / conv
  PK_OTHER = 9,
  PK_EDGE_CHAIN = 10,  // fused tcgen05 edge stage (gather -> 3 UMMA layers -> LN -> segmented reduce)
  PK_EDGE_CHAIN_BWD = 11,  // fused tcgen05 backward of the edge stage
  PK_COUNT = 12
};
bool prof_enabled();
void prof_begin(int kind, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(int kind, cudaStream_t s) : st(s), on(prof_enabled()) {
    if (on) prof_begin(kind, st);
  }
  ~ProfScope() {
    if (on) prof_end(st);
  }
};

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace bsms
