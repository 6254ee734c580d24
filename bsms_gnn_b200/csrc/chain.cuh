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

// Row-cooperative gather of the first edge activation a0 = relu(Ps[src] + Pd[dst] + F fiber) into a
// K-major SWIZZLE_128B bf16 operand tile ([row][channel], the layout of wblk_offset): ONE WARP PER
// ROW, lane l owns channels 4l..4l+3, so every global load is a coalesced 512 B row (4 L1 wavefronts
// instead of the 32 a lane-per-row gather costs) and U rows (2U loads) are in flight per warp.
// s_ij[r] = (b*N+src, b*N+dst) of tile row r, (-1, -1) for rows past the end (their fiber is 0 and
// the tile row is zero-filled); Fl[c] = fiber coefficients of channel 4l+c.
template <int U>
__device__ __forceinline__ void coop_gather_a0(const float* __restrict__ PsPd, const int2* s_ij, const float4* s_fib,
                                               const float4 (&Fl)[4], uint8_t* tile, int r_begin, int r_end, int lane,
                                               float* dbg, long long dbg_row0) {
  const uint32_t col_off = (uint32_t)((lane >> 4) * 16384 + (lane & 1) * 8);
  const int chunk7 = (lane >> 1) & 7;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int r = r_begin; r < r_end; r += U) {
    float4 a[U], d[U];
    int ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int2 ij = s_ij[r + u];
      ok[u] = ij.x >= 0;
      a[u] = ok[u] ? ld4(PsPd + (size_t)ij.x * 256 + 4 * lane) : z4;
      d[u] = ok[u] ? ld4(PsPd + (size_t)ij.y * 256 + 128 + 4 * lane) : z4;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 f = s_fib[r + u];
      float x0 = a[u].x + d[u].x, x1 = a[u].y + d[u].y, x2 = a[u].z + d[u].z, x3 = a[u].w + d[u].w;
      x0 = x0 + Fl[0].x * f.x + Fl[0].y * f.y + Fl[0].z * f.z + Fl[0].w * f.w;
      x1 = x1 + Fl[1].x * f.x + Fl[1].y * f.y + Fl[1].z * f.z + Fl[1].w * f.w;
      x2 = x2 + Fl[2].x * f.x + Fl[2].y * f.y + Fl[2].z * f.z + Fl[2].w * f.w;
      x3 = x3 + Fl[3].x * f.x + Fl[3].y * f.y + Fl[3].z * f.z + Fl[3].w * f.w;
      x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
      if (dbg && ok[u]) st4(dbg + (dbg_row0 + r + u) * 128 + 4 * lane, make_float4(x0, x1, x2, x3));
      uint2 pk;
      pk.x = pack_bf16(x0, x1);
      pk.y = pack_bf16(x2, x3);
      *reinterpret_cast<uint2*>(tile + col_off + (r + u) * 128 + ((chunk7 ^ ((r + u) & 7)) << 4)) = pk;
    }
  }
}

// The same gather, software-pipelined: the loads of the next U rows are issued before the current U rows
// are combined, so a warp that gathers alone on its scheduler (warp-specialised producer) keeps 2U
// row requests in flight instead of stalling on every batch.  (r_end - r_begin) must be a multiple of U.
template <int U>
__device__ __forceinline__ void coop_gather_a0_pipe(const float* __restrict__ PsPd, const int2* s_ij, const float4* s_fib,
                                                    const float4 (&Fl)[4], uint8_t* tile, int r_begin, int r_end, int lane,
                                                    float* dbg, long long dbg_row0) {
  const uint32_t col_off = (uint32_t)((lane >> 4) * 16384 + (lane & 1) * 8);
  const int chunk7 = (lane >> 1) & 7;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a[2][U], d[2][U];
  auto issue = [&](int r, float4 (&aa)[U], float4 (&dd)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int2 ij = s_ij[r + u];
      const bool ok = ij.x >= 0;
      aa[u] = ok ? ld4(PsPd + (size_t)ij.x * 256 + 4 * lane) : z4;
      dd[u] = ok ? ld4(PsPd + (size_t)ij.y * 256 + 128 + 4 * lane) : z4;
    }
  };
  auto combine = [&](int r, const float4 (&aa)[U], const float4 (&dd)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 f = s_fib[r + u];
      float x0 = aa[u].x + dd[u].x, x1 = aa[u].y + dd[u].y, x2 = aa[u].z + dd[u].z, x3 = aa[u].w + dd[u].w;
      x0 = x0 + Fl[0].x * f.x + Fl[0].y * f.y + Fl[0].z * f.z + Fl[0].w * f.w;
      x1 = x1 + Fl[1].x * f.x + Fl[1].y * f.y + Fl[1].z * f.z + Fl[1].w * f.w;
      x2 = x2 + Fl[2].x * f.x + Fl[2].y * f.y + Fl[2].z * f.z + Fl[2].w * f.w;
      x3 = x3 + Fl[3].x * f.x + Fl[3].y * f.y + Fl[3].z * f.z + Fl[3].w * f.w;
      x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
      if (dbg && s_ij[r + u].x >= 0) st4(dbg + (dbg_row0 + r + u) * 128 + 4 * lane, make_float4(x0, x1, x2, x3));
      uint2 pk;
      pk.x = pack_bf16(x0, x1);
      pk.y = pack_bf16(x2, x3);
      *reinterpret_cast<uint2*>(tile + col_off + (r + u) * 128 + ((chunk7 ^ ((r + u) & 7)) << 4)) = pk;
    }
  };
  issue(r_begin, a[0], d[0]);
#pragma unroll 1
  for (int r = r_begin; r < r_end; r += 2 * U) {
    if (r + U < r_end) issue(r + U, a[1], d[1]);
    combine(r, a[0], d[0]);
    if (r + U < r_end) {
      if (r + 2 * U < r_end) issue(r + 2 * U, a[0], d[0]);
      combine(r + U, a[1], d[1]);
    }
  }
}

// Row-cooperative load of NR contiguous-in-index rows (tile rows r_begin .. r_begin+NR-1 = global rows
// row0 + r) into registers (lane l: channels 4l..4l+3 of each row; rows past `rows` read as zero), and
// their conversion into a K-major SWIZZLE_128B bf16 operand tile.  Split so that the loads of the NEXT
// tile can be in flight while the current tile computes.
template <int NR>
__device__ __forceinline__ void coop_rows_load(const float* __restrict__ X, int ld, long long row0, long long rows,
                                               int r_begin, int lane, float4 (&v)[NR]) {
#pragma unroll
  for (int u = 0; u < NR; ++u) {
    const long long row = row0 + r_begin + u;
    v[u] = row < rows ? ld4(X + row * ld + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int NR>
__device__ __forceinline__ void coop_rows_store(uint8_t* tile, int r_begin, int lane, const float4 (&v)[NR]) {
  const uint32_t col_off = (uint32_t)((lane >> 4) * 16384 + (lane & 1) * 8);
  const int chunk7 = (lane >> 1) & 7;
#pragma unroll
  for (int u = 0; u < NR; ++u) {
    const int r = r_begin + u;
    uint2 pk;
    pk.x = pack_bf16(v[u].x, v[u].y);
    pk.y = pack_bf16(v[u].z, v[u].w);
    *reinterpret_cast<uint2*>(tile + col_off + r * 128 + ((chunk7 ^ (r & 7)) << 4)) = pk;
  }
}

// the same rows as a two-way fp16 split of v * scale: hi = fp16(v s) -> tile_hi, lo = fp16(v s - hi) -> tile_lo (22
// significant bits; s is a power of two; values are clamped to the fp16 range so an outlier saturates instead of
// becoming infinite)
template <int NR>
__device__ __forceinline__ void coop_rows_store_split(uint8_t* tile_hi, uint8_t* tile_lo, int r_begin, int lane, const float4 (&v)[NR],
                                                      float scale) {
  const uint32_t col_off = (uint32_t)((lane >> 4) * 16384 + (lane & 1) * 8);
  const int chunk7 = (lane >> 1) & 7;
  auto cl = [](float x) { return fminf(fmaxf(x, -65000.f), 65000.f); };
#pragma unroll
  for (int u = 0; u < NR; ++u) {
    const int r = r_begin + u;
    const float s0 = cl(v[u].x * scale), s1 = cl(v[u].y * scale), s2 = cl(v[u].z * scale), s3 = cl(v[u].w * scale);
    const __half h0 = __float2half_rn(s0), h1 = __float2half_rn(s1), h2 = __float2half_rn(s2), h3 = __float2half_rn(s3);
    uint2 ph, pl;
    ph.x = pack_f16(h0, h1);
    ph.y = pack_f16(h2, h3);
    pl.x = pack_f16(__float2half_rn(s0 - __half2float(h0)), __float2half_rn(s1 - __half2float(h1)));
    pl.y = pack_f16(__float2half_rn(s2 - __half2float(h2)), __float2half_rn(s3 - __half2float(h3)));
    const uint32_t off = col_off + r * 128 + ((chunk7 ^ (r & 7)) << 4);
    *reinterpret_cast<uint2*>(tile_hi + off) = ph;
    *reinterpret_cast<uint2*>(tile_lo + off) = pl;
  }
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


// ---- deterministic option (bsms_set_deterministic) in the bf16 mode: every kernel that finishes with one atomic
// flush per CTA writes its per-CTA (per-warp) partial sums to a scratch block instead, and one small kernel adds
// them up in a fixed order (tiles are assigned to CTAs statically, so a CTA's partial is itself reproducible); the
// two fused edge kernels write their per-edge-row results as rows and order-fixed CSR segment sums replace the
// red.add reductions into node rows.
int det_enabled();  // gmp.cu
constexpr int kDetEdgeBwdStride = 3 * 16384 + 3 * 8 * 128 + 8 * 128 + 8 * 128 * 4;  // gW2..4 | gb2..4 per warp | gb1 per warp | fiber per warp
constexpr int kDetNodeBwdStride = 3 * 16384 + 3 * 8 * 128;                           // gV2..4 | gc2..4 per warp
constexpr int kDetWgradStride = 16384 + 2 * 128;                                     // dW | db per thread half
struct DetSeg {
  float* dst;  // dst[row * ldd + col] += sum over parts part0..part0+nparts-1 and reps of src[rep][row][col]
  int src_off, part0, nparts, reps, rows, cols, src_cols, ldd;
};
int det_reduce(const float* part, int stride, const DetSeg* segs, int nseg, cudaStream_t st);  // gmp_tc.cu
size_t det_part_bytes();                                                                       // gmp_tc.cu
}  // namespace bsms
