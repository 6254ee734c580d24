// Shared pieces of the fused tcgen05 MLP-chain kernels: packed-operand layout and the weight packer.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace bsms {
using namespace umma;

constexpr int kD = BSMS_LATENT;
constexpr uint32_t kWBlk = 128 * 128 * 2;  // one packed 128x128 16-bit operand block (32 KB)
constexpr float kActScale = 16.f;          // 2^4
constexpr float kWScale = 256.f;           // 2^8

// byte offset of element (n, k) inside a packed block = its shared-memory image
__host__ __device__ inline uint32_t wblk_offset(int n, int k) {
  return (uint32_t)((k >> 6) * 16384 + n * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + (k & 7) * 2);
}

struct PackList {
  const float* w[12];
  int ld[12];
  int n;
};

// fp32 [128 x 128] (row n, ld) -> packed 16-bit block(s).  grid = (n_blocks), block = 256
template <int NSPLIT>
__global__ void k_pack_weights(PackList pl, uint8_t* __restrict__ out) {
  const int blk = blockIdx.x;
  const float* W = pl.w[blk];
  const int ld = pl.ld[blk];
  uint8_t* o = out + (size_t)blk * NSPLIT * kWBlk;
  for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x) {
    int n = idx >> 7, k = idx & 127;
    float v = W[(size_t)n * ld + k];
    uint32_t off = wblk_offset(n, k);
    if (NSPLIT == 1) {
      *reinterpret_cast<__nv_bfloat16*>(o + off) = __float2bfloat16_rn(v);
    } else {
      float s = v * kWScale;
      __half hi = __float2half_rn(s);
      __half lo = __float2half_rn(s - __half2float(hi));
      *reinterpret_cast<__half*>(o + off) = hi;
      *reinterpret_cast<__half*>(o + kWBlk + off) = lo;
    }
  }
}


// one weight-gradient problem dW[128, ldo] += G^T X over `rows` rows (node_gemm.cu)
struct WgradParams {
  const float* G;
  int ldg;
  const float* X;
  int ldx;
  float* dW;  // [128, ldo] accumulated
  int ldo;
  float* db;  // [128] accumulated, may be null
  long long rows;
  int ntiles;
};

}  // namespace bsms
