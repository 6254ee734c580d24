// Fused backward of the GMP edge stage on tcgen05 (bf16 operands, fp32 accumulate), by
// recomputation: nothing per-edge is read from or written to HBM except the gathers of the
// projected node rows / upstream gradient rows and the scatter of the edge-input gradient.
//
// Per 128-edge tile (one tile per CTA at a time, persistent grid, 256 threads = 2 threads per edge
// row: thread (row r, half h) owns channels [64h, 64h+64) of TMEM lane r):
//   a0 = relu(Ps[src]+Pd[dst]+b1+F fiber)            row-cooperative gather (one warp per 512 B row) -> T0 (bf16 smem tile)
//   a1 = relu(a0 W2^T + b2)                           UMMA  D=T0 x W2   -> T1
//   a2 = relu(a1 W3^T + b3)                           UMMA  D=T1 x W3   -> T2
//   y  = a2 W4^T + b4 ; gy = LN'(y) * g_aggr[dst]     UMMA  D=T2 x W4   -> the W2 slot (W2 is idle until the last dgrad;
//                                                     cp.async.bulk brings it back from L2 once gy is consumed)
//   g2 = (gy W4) . [a2>0] ; dW4 += gy^T a2            UMMA  dgrad D=gy x W4(MN) -> T2 (in place of a2), wgrad(gy,T2) behind it
//   g1 = (g2 W3) . [a1>0] ; dW3 += g2^T a1            UMMA  dgrad D=T2 x W3(MN) -> T1 (in place of a1), wgrad(T2,T1)
//   g0 = (g1 W2) . [a0>0] ; dW2 += g1^T a0            UMMA  dgrad D=T1 x W2(MN) -> fp32 staging over T1|T2, wgrad(T1,T0)
//   gPs[src] += g0 (one coalesced 512 B red.add.v4 per row) ; gPd[dst] += run sums of g0 ; gF += g0^T fiber ; gb* += column sums
// The three weight-gradient accumulators (3 x 128 TMEM columns) stay resident in tensor memory for
// the whole persistent loop and are reduced into global memory once per CTA; the activation /
// gradient tiles are written once in the canonical 128B-swizzled layout and consumed BOTH as
// K-major operands (forward / data-gradient GEMMs) and as MN-major operands (weight-gradient
// GEMMs) by re-describing the same bytes.  Weights are staged once per CTA by cp.async.bulk and
// read K-major (recompute) and MN-major (= W^T, data gradient) from the same image.
#include <stdlib.h>

#include "chain.cuh"

namespace bsms {

struct EdgeBwdParams {
  const float* PsPd;  // [B*N, 256], b1 folded into the Pd half
  const float* pos;
  int pos_batched, P;
  const int32_t* src_d;
  const int32_t* dst_d;
  const float* W1;
  const float* b[4];
  const uint8_t* wpack;  // [3] packed bf16 blocks (W2, W3, W4)
  const float* g_aggr;   // upstream gradient of aggr, row stride ld_g
  int ld_g;
  float* gPsPd;  // [B*N, 256], zero-initialised; receives gPs | gPd
  float* g0_rows;  // deterministic variant: the edge-input gradient leaves as rows [B*E, 128] (no scatter) ...
  float* part;     // ... and the per-CTA sums go to part[blockIdx.x][kDetEdgeBwdStride] (no atomic flush)
  float* gW[3];  // W2, W3, W4 gradients [128,128] (accumulated)
  float* gb[4];  // b1..b4 gradients
  float* gW1;    // mlp_edge layer-0 weight gradient [128, 2*128+P+1]: fiber columns accumulated here
  int B, N, E;
  long long rows;
  int ntiles;
  unsigned long long* prof;  // optional [16] per-phase cycle counters (BSMS_PHASE_PROF=1), summed over CTAs
};

__device__ __forceinline__ uint32_t tile_off(int r, int chunk) {  // 16-byte chunk `chunk` (0..15) of row r
  return (uint32_t)((chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4));
}

__device__ __forceinline__ void load_d64(uint32_t taddr, float (&v)[64]) {
  uint32_t r0[32], r1[32];
  tmem_ld32(taddr, r0);
  tmem_ld32(taddr + 32, r1);
  wait_ld();
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    v[t] = __uint_as_float(r0[t]);
    v[32 + t] = __uint_as_float(r1[t]);
  }
}

// fiber = [pos_i - pos_j, |pos_i - pos_j|] padded to 4 (src/ops/basic.py:83-85), P in {1,2,3} at run time, scalars
// only: a local array indexed by P would live in local memory and every use would wait for its loads
__device__ __forceinline__ float4 make_fiber(const float* __restrict__ pb, int P, int i, int j) {
  const float* a = pb + (size_t)i * P;
  const float* b = pb + (size_t)j * P;
  const float d0 = a[0] - b[0];
  const float d1 = P > 1 ? a[1] - b[1] : 0.f;
  const float d2 = P > 2 ? a[2] - b[2] : 0.f;
  const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
  return P == 1 ? make_float4(d0, nrm, 0.f, 0.f) : (P == 2 ? make_float4(d0, d1, nrm, 0.f) : make_float4(d0, d1, d2, nrm));
}

// Structure notes (measured on B200, profiles/r2_phase_cycles.txt, profiles/r2_edge_kernels_ncu.txt):
//  * ReLU + bf16 packing is one cvt (F2FP.RELU), ReLU masks are re-derived from the stored activation tiles
//    (a > 0 <=> its bf16 image is non-zero), LayerNorm backward is two FMAs per element, and the
//    weight-gradient MMA of every backward layer is issued AFTER its data-gradient MMA on its own barrier,
//    so it runs under the epilogue of the data gradient instead of in front of it (8.66 -> 7.88 ms / step).
//  * Measured and rejected: N-split GEMM groups with half-epilogues under the second half's MMA
//    (8.42 vs 7.91 ms: two N = 64 groups re-read the A operand and take longer than one N = 128 group); the
//    sender-side scatter as 512-byte bulk reductions (UBLKRED, 8.16 vs 7.88 ms: ~30 cycles of issue each); the
//    gather of tile i+1 interleaved with the scatter rows of tile i (7.75 vs 7.57 ms: loads and reductions share
//    the SM's L1/L2 port, the merged phase costs the sum of the two); rows 0..63 of the next tile gathered into a
//    16 KB side buffer inside the shadow of the first data-gradient GEMM (7.75 vs 7.54 ms: the tile start gets
//    1.3 k cycles shorter, the shadow 1.45 k longer — an MMA pair lasts ~1 k cycles, a batch of gathers ~2 k).
//  * F2: the epilogue arithmetic uses the packed fp32x2 instructions of sm_100 (FADD2 / FFMA2): the epilogues
//    are bound by the FMA pipe's issue rate, not by latency.
template <bool PROF, bool F2, bool DET>
__global__ void __launch_bounds__(256, 1) k_edge_chain_bwd(const EdgeBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr bool NS = false;  // N-split GEMM groups (two N = 64 halves on their own barriers): measured slower, 8.42 vs 7.91 ms
  constexpr uint32_t NH = NS ? 64 : 128;                          // N of one MMA group
  constexpr uint32_t IDESC_KK = make_idesc(1, 128, NH, 0, 0);     // A K-major, B K-major   (recompute)
  constexpr uint32_t IDESC_KM = make_idesc(1, 128, NH, 0, 1);     // A K-major, B MN-major  (dgrad: B = W^T)
  constexpr uint32_t IDESC_MM = make_idesc(1, 128, 128, 1, 1);    // A, B MN-major          (wgrad)
  const uint32_t s0 = smem_u32(smem_raw);
  const uint32_t sbase = (s0 + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sbase - s0);
  uint8_t* s_T[3] = {sp + 3 * kWBlk, sp + 4 * kWBlk, sp + 5 * kWBlk};
  float* s_bias = reinterpret_cast<float*>(sp + 6 * kWBlk);  // [4][128]
  float4* s_F = reinterpret_cast<float4*>(s_bias + 512);     // [128]
  float4* s_fib2 = s_F + 128;                                // [2][128] fiber of each tile row (double-buffered)
  float4* s_x = s_fib2 + 256;                                // [2][128] LayerNorm partial sums
  int2* s_ij2 = reinterpret_cast<int2*>(s_x + 256);          // [2][128] (b*N+src, b*N+dst) of each tile row, -1 past the end
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_ij2 + 256);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 5);
  const uint32_t aW[3] = {sbase, sbase + kWBlk, sbase + 2 * kWBlk};
  const uint32_t aT[3] = {sbase + 3 * kWBlk, sbase + 4 * kWBlk, sbase + 5 * kWBlk};

  const int tid = threadIdx.x, warp = (int)uniform(threadIdx.x >> 5), lane = tid & 31;
  const int q = warp & 3, h = warp >> 2, r = q * 32 + lane;
  const uint32_t bar_w = smem_u32(&s_bar[0]), bar_m = smem_u32(&s_bar[1]), bar_r = smem_u32(&s_bar[2]);
  const uint32_t bar_g = smem_u32(&s_bar[3]);  // completion of the weight-gradient MMA issued behind a data-gradient MMA
  const uint32_t bar_h = smem_u32(&s_bar[4]);  // NS: completion of the second N half

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    mbar_init(bar_r, 1);
    mbar_init(bar_g, 1);
    mbar_init(bar_h, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 512);
  // b2..b4 are rounded to bf16 exactly as the forward kernel sees them (they ride in its MMA)
  for (int i = tid; i < 512; i += 256) {
    const float bv = p.b[i >> 7][i & 127];
    s_bias[i] = i < 128 ? bv : __bfloat162float(__float2bfloat16_rn(bv));
  }
  {
    const int ldw1 = 2 * kD + p.P + 1;
    for (int c = tid; c < 128; c += 256) {
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k <= p.P; ++k) f[k] = p.W1[(size_t)c * ldw1 + k];
      s_F[c] = make_float4(f[0], f[1], f[2], f[3]);
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = uniform(*s_tmem);
  if (tid == 0) {
    mbar_expect_tx(bar_w, 3 * kWBlk);
    for (int blk = 0; blk < 3; ++blk) bulk_g2s(aW[blk], p.wpack + (size_t)blk * kWBlk, kWBlk, bar_w);
    mbar_wait(bar_w, 0);
  }
  __syncwarp();  // lane 0 rejoins its warp (a warp left split runs its collectives on the slow path until the next barrier)
  const uint32_t d_tmem = tmem_base;  // D: cols [0,128); dW2/dW3/dW4: cols [128,256), [256,384), [384,512)
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  // this thread's two groups of 32 channels (= TMEM columns of D): NS: one group in each N half
  const int ch0[2] = {NS ? 32 * h : 64 * h, NS ? 64 + 32 * h : 64 * h + 32};
  const uint32_t d_col[2] = {d_tmem + lane_off + (uint32_t)ch0[0], d_tmem + lane_off + (uint32_t)ch0[1]};
  uint32_t phase = 0, phase_r = 0, phase_g = 0, phase_h = 0;
  uint8_t* s_gy = sp;  // the W2 slot: W2 is idle between the first recompute GEMM and the last data-gradient GEMM,
                       // so gy lives there meanwhile and W2 is brought back by cp.async.bulk (32 KB from L2 per tile)
  uint32_t wacc = 0;  // weight-gradient accumulators hold something
  // optional phase profile: thread 0 accumulates the cycles between consecutive marks
  unsigned long long* s_prof = reinterpret_cast<unsigned long long*>(s_tmem + 2);  // [16]
  long long tprev = 0;
  if (PROF && tid == 0) {
    for (int k = 0; k < 16; ++k) s_prof[k] = 0ull;
    tprev = clock64();
  }
  auto mark = [&](int k) {
    if (PROF && tid == 0) {
      const long long t = clock64();
      s_prof[k] += (unsigned long long)(t - tprev);
      tprev = t;
    }
  };
  // persistent per-thread partial sums for the bias / fiber-weight gradients: lane l owns channels 4l..4l+3
  float4 acc_b[4];  // layers 2..4 ([0] unused: acc_b0 below)
#pragma unroll
  for (int l = 0; l < 4; ++l) acc_b[l] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc_b0 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc_fl[4];  // [fiber component k] x 4 channels
#pragma unroll
  for (int k = 0; k < 4; ++k) acc_fl[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float* s_g0 = reinterpret_cast<float*>(s_T[1]);  // fp32 [128][128] staging of g0 over T1|T2 (64 KB)

  // one full-CTA phase boundary: make generic smem writes visible to the tensor core, order tcgen05 ops
  auto sync_all = [&]() {
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  };
  // the accumulator columns of channel group hh are complete (NS: group hh lives in N half hh)
  auto wait_half = [&](int hh) {
    if (hh == 0) {
      mbar_wait(bar_m, phase);
      phase ^= 1;
      fence_after_sync();
    } else if (NS) {
      mbar_wait(bar_h, phase_h);
      phase_h ^= 1;
      fence_after_sync();
    }
  };
  // D = A(tile, K-major) x B(weight block): K-major B for the recompute, MN-major B (= W^T) for dgrad;
  // NS: two N = 64 groups, each committed to its own barrier
  auto issue_gemm = [&](uint32_t a_tile, uint32_t b_blk, bool b_mn) {
#pragma unroll
    for (int hn = 0; hn < (NS ? 2 : 1); ++hn) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t ad = smem_desc_sw128(a_tile + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
        // K-major B: rows n of the [n][k] image, N half = +64 rows; MN-major B: N runs along the 128-byte atom,
        // N half = the next atom (+16 KB)
        const uint64_t bd = b_mn ? smem_desc_sw128(b_blk + hn * 16384 + ks * 2048, 16384, 1024)
                                 : smem_desc_sw128(b_blk + (ks >> 2) * 16384 + hn * 8192 + (ks & 3) * 32, 16, 1024);
        mma_ss(d_tmem + 64 * hn, ad, bd, b_mn ? IDESC_KM : IDESC_KK, ks > 0);
      }
      mma_commit(hn == 0 ? bar_m : bar_h);
    }
  };
  // dW[out][in] += G^T A : both operands MN-major views of [row][channel] tiles, K = 128 tile rows
  auto issue_wgrad = [&](uint32_t dw_tmem, uint32_t g_tile, uint32_t a_tile) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t ad = smem_desc_sw128(g_tile + ks * 2048, 16384, 1024);
      const uint64_t bd = smem_desc_sw128(a_tile + ks * 2048, 16384, 1024);
      mma_ss(dw_tmem, ad, bd, IDESC_MM, (wacc | ks) != 0);
    }
    mma_commit(bar_g);
  };
  // column sums of a gradient tile (bias gradient), row-cooperative: warp w sums rows 16w..16w+15, lane l
  // owns channels 4l..4l+3 (one 8-byte shared load per row); overlaps the MMAs
  auto colsum = [&](const uint8_t* tile, float4& acc) {
    const uint8_t* base = tile + (lane >> 4) * 16384 + (lane & 1) * 8;
    const int chunk7 = (lane >> 1) & 7;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int rr = warp * 16; rr < warp * 16 + 16; ++rr) {
      const uint2 u = *reinterpret_cast<const uint2*>(base + rr * 128 + ((chunk7 ^ (rr & 7)) << 4));
      s.x += __uint_as_float(u.x << 16); s.y += __uint_as_float(u.x & 0xFFFF0000u);
      s.z += __uint_as_float(u.y << 16); s.w += __uint_as_float(u.y & 0xFFFF0000u);
    }
    acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w;
  };
  auto gather_a0 = [&](const int2* s_ij, const float4* s_fib) {
    float4 Fl[4];  // fiber coefficients of this lane's 4 channels
#pragma unroll
    for (int c = 0; c < 4; ++c) Fl[c] = s_F[4 * lane + c];
    coop_gather_a0<8>(p.PsPd, s_ij, s_fib, Fl, s_T[0], warp * 16, warp * 16 + 16, lane, nullptr, 0);
  };
  // (D + bias) of 8 consecutive accumulator columns -> fp32 pairs
  auto add_bias8 = [&](const uint32_t* rr8, const float4 ba, const float4 bb, float2 (&x)[4]) {
    if constexpr (F2) {
      x[0] = __fadd2_rn(make_float2(__uint_as_float(rr8[0]), __uint_as_float(rr8[1])), make_float2(ba.x, ba.y));
      x[1] = __fadd2_rn(make_float2(__uint_as_float(rr8[2]), __uint_as_float(rr8[3])), make_float2(ba.z, ba.w));
      x[2] = __fadd2_rn(make_float2(__uint_as_float(rr8[4]), __uint_as_float(rr8[5])), make_float2(bb.x, bb.y));
      x[3] = __fadd2_rn(make_float2(__uint_as_float(rr8[6]), __uint_as_float(rr8[7])), make_float2(bb.z, bb.w));
    } else {
      x[0] = make_float2(__uint_as_float(rr8[0]) + ba.x, __uint_as_float(rr8[1]) + ba.y);
      x[1] = make_float2(__uint_as_float(rr8[2]) + ba.z, __uint_as_float(rr8[3]) + ba.w);
      x[2] = make_float2(__uint_as_float(rr8[4]) + bb.x, __uint_as_float(rr8[5]) + bb.y);
      x[3] = make_float2(__uint_as_float(rr8[6]) + bb.z, __uint_as_float(rr8[7]) + bb.w);
    }
  };

  int it = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    // ---- row metadata, double-buffered (first 128 threads: one tile row each).  Only the first tile computes
    //      it here; the metadata of every later tile is produced in three stages inside the recompute GEMM
    //      waits of the tile before it (index loads / position loads / fiber), where every thread idles anyway.
    float4* s_fib = s_fib2 + (it & 1) * 128;
    int2* s_ij = s_ij2 + (it & 1) * 128;
    if (it == 0 && tid < 128) {
      const long long row_ = (long long)tile * 128 + tid;
      float4 fib4 = make_float4(0.f, 0.f, 0.f, 0.f);
      int2 ij = make_int2(-1, -1);
      if (row_ < p.rows) {
        const int b_ = (int)(row_ / p.E);
        const int e_ = (int)(row_ - (long long)b_ * p.E);
        const int i_ = p.src_d[e_], j_ = p.dst_d[e_];
        fib4 = make_fiber(p.pos + (p.pos_batched ? (size_t)b_ * p.N * p.P : 0), p.P, i_, j_);
        ij = make_int2(b_ * p.N + i_, b_ * p.N + j_);
      }
      s_fib[tid] = fib4;
      s_ij[tid] = ij;
    }
    __syncthreads();
    mark(0);
    const long long row = (long long)tile * 128 + r;
    const bool valid = row < p.rows;
    const int rowj = s_ij[r].y;

    // activation epilogue: relu(D + bias) -> tile; ReLU rides in the bf16 conversion, no mask is kept
    auto act_epilogue = [&](const float* bias, uint8_t* dst_tile) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        wait_half(hh);
        uint32_t rr_[32];
        tmem_ld32(d_col[hh], rr_);
        wait_ld();
        const float4* b4 = reinterpret_cast<const float4*>(bias + ch0[hh]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float2 x[4];
          add_bias8(rr_ + 8 * jj, b4[2 * jj], b4[2 * jj + 1], x);
          uint4 u;
          u.x = pack_relu_bf16(x[0].x, x[0].y);
          u.y = pack_relu_bf16(x[1].x, x[1].y);
          u.z = pack_relu_bf16(x[2].x, x[2].y);
          u.w = pack_relu_bf16(x[3].x, x[3].y);
          *reinterpret_cast<uint4*>(dst_tile + tile_off(r, (ch0[hh] >> 3) + jj)) = u;
        }
      }
    };
    // gradient epilogue: g = D . [act > 0] written IN PLACE over the activation tile `tile` (same thread, same
    // bytes).  The weight-gradient MMA issued behind the data-gradient MMA still reads `tile` (and the gradient
    // tile before it): the TMEM load and the arithmetic of the first group run under it, the first store waits
    // for its barrier.
    auto grad_epilogue = [&](uint8_t* tile) {
      const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        wait_half(hh);
        uint32_t rr_[32];
        tmem_ld32(d_col[hh], rr_);
        wait_ld();
        uint4 u[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint4 a8 = *reinterpret_cast<const uint4*>(tile + tile_off(r, (ch0[hh] >> 3) + jj));
          const uint32_t aw[4] = {a8.x, a8.y, a8.z, a8.w};
          uint32_t o[4];
#pragma unroll
          for (int w2 = 0; w2 < 4; ++w2) {
            const uint32_t gp = pack_bf16(__uint_as_float(rr_[8 * jj + 2 * w2]), __uint_as_float(rr_[8 * jj + 2 * w2 + 1]));
            const __nv_bfloat162 on = __hgt2(*reinterpret_cast<const __nv_bfloat162*>(&aw[w2]), z2);  // 1.0 / 0.0 per half
            const __nv_bfloat162 gm = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&gp), on);
            o[w2] = *reinterpret_cast<const uint32_t*>(&gm);
          }
          u[jj] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        if (hh == 0) {  // the weight-gradient MMA behind this data gradient has read the tile
          mbar_wait(bar_g, phase_g);
          phase_g ^= 1;
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) *reinterpret_cast<uint4*>(tile + tile_off(r, (ch0[hh] >> 3) + jj)) = u[jj];
      }
    };

    // ---- recompute the forward chain
    gather_a0(s_ij, s_fib);
    sync_all();
    mark(1);
    if (warp == 0) {  // one elected lane issues; operands are warp-uniform (no per-lane R2UR loop per MMA)
      if (elect_one()) issue_gemm(aT[0], aW[0], false);
      __syncwarp();
    }
    // in the shadow of GEMM 1: index loads of the next tile's rows (stage 1) and the L2 prefetch of its rows
    int2 nij = make_int2(-1, -1);
    int nbatch = 0;
    if (tid < 128) {
      const long long nrow_ = (long long)(tile + gridDim.x) * 128 + tid;
      if (tile + (int)gridDim.x < p.ntiles && nrow_ < p.rows) {
        nbatch = (int)(nrow_ / p.E);
        const int e_ = (int)(nrow_ - (long long)nbatch * p.E);
        nij = make_int2(p.src_d[e_], p.dst_d[e_]);
      }
    }
    {
      if (nij.x >= 0) {
        const float* nps = p.PsPd + ((size_t)nbatch * p.N + nij.x) * 256;
        const float* npd = p.PsPd + ((size_t)nbatch * p.N + nij.y) * 256 + 128;
        const float* ng = p.g_aggr + ((size_t)nbatch * p.N + nij.y) * p.ld_g;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nps + k * 32));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(npd + k * 32));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ng + k * 32));
        }
      }
    }
    mark(2);
    act_epilogue(s_bias + 128, s_T[1]);
    sync_all();
    mark(3);
    if (warp == 0) {
      if (elect_one()) issue_gemm(aT[1], aW[1], false);
      __syncwarp();
    }
    // in the shadow of GEMM 2: positions of the next tile's end points -> fiber (stage 2; plain scalars, the loads are
    // consumed in the shadow of GEMM 3)
    float4 nfib = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nij.x >= 0) nfib = make_fiber(p.pos + (p.pos_batched ? (size_t)nbatch * p.N * p.P : 0), p.P, nij.x, nij.y);
    mark(4);
    act_epilogue(s_bias + 256, s_T[2]);
    sync_all();
    mark(5);
    if (warp == 0) {
      if (elect_one()) issue_gemm(aT[2], aW[2], false);
      __syncwarp();
    }
    if (tid < 128) {  // in the shadow of GEMM 3: the next tile's row metadata -> the other buffer (stage 3)
      s_fib2[((it + 1) & 1) * 128 + tid] = nfib;
      s_ij2[((it + 1) & 1) * 128 + tid] = nij.x >= 0 ? make_int2(nbatch * p.N + nij.x, nbatch * p.N + nij.y) : make_int2(-1, -1);
    }
    // upstream gradient row g_aggr[dst] (this thread's 2 x 32 channels): issued before the MMA wait
    float2 g[32];
    {
      const float* grow = p.g_aggr + (size_t)(valid ? rowj : 0) * p.ld_g;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 gv = valid ? ld4(grow + ch0[hh] + q4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          g[16 * hh + 2 * q4] = make_float2(gv.x, gv.y);
          g[16 * hh + 2 * q4 + 1] = make_float2(gv.z, gv.w);
        }
      }
    }
    mark(6);
    // ---- LayerNorm backward: gy = rstd * (g - mean(g) - yhat * mean(g * yhat)) -> the W2 slot
    //      two sweeps over this thread's 64 accumulator columns, 32 at a time (register budget)
    {
      float2 t1 = make_float2(0.f, 0.f), t2 = t1, t3 = t1, t4 = t1;  // even / odd partial sums of y, y^2, g, g y
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        wait_half(hh);
        uint32_t rr_[32];
        tmem_ld32(d_col[hh], rr_);
        wait_ld();
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + 384 + ch0[hh]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float2 y[4];
          add_bias8(rr_ + 8 * jj, b4[2 * jj], b4[2 * jj + 1], y);
#pragma unroll
          for (int w2 = 0; w2 < 4; ++w2) {
            const float2 gp = g[16 * hh + 4 * jj + w2];
            if constexpr (F2) {
              t1 = __fadd2_rn(t1, y[w2]);
              t2 = __ffma2_rn(y[w2], y[w2], t2);
              t3 = __fadd2_rn(t3, gp);
              t4 = __ffma2_rn(gp, y[w2], t4);
            } else {
              t1.x += y[w2].x; t1.y += y[w2].y;
              t2.x = fmaf(y[w2].x, y[w2].x, t2.x); t2.y = fmaf(y[w2].y, y[w2].y, t2.y);
              t3.x += gp.x; t3.y += gp.y;
              t4.x = fmaf(gp.x, y[w2].x, t4.x); t4.y = fmaf(gp.y, y[w2].y, t4.y);
            }
          }
        }
      }
      float s1 = t1.x + t1.y, s2 = t2.x + t2.y, s3 = t3.x + t3.y, s4 = t4.x + t4.y;
      s_x[h * 128 + r] = make_float4(s1, s2, s3, s4);
      __syncthreads();
      const float4 o = s_x[(1 - h) * 128 + r];
      s1 += o.x; s2 += o.y; s3 += o.z; s4 += o.w;
      const float mean = s1 * (1.f / 128.f);
      const float var = fmaxf(s2 * (1.f / 128.f) - mean * mean, 0.f);
      const float rstd = 1.f / sqrtf(var + 1e-5f);
      const float c1 = s3 * (1.f / 128.f);
      const float c2 = rstd * (s4 - mean * s3) * (1.f / 128.f);
      // rstd (g - c1 - yh c2) with yh = (y - mean) rstd  ==  kA g + kB y + kC  (two FMAs per element)
      const float kA = valid ? rstd : 0.f;
      const float kB = valid ? -rstd * rstd * c2 : 0.f;
      const float kC = valid ? rstd * (rstd * c2 * mean - c1) : 0.f;
      const float2 kA2 = make_float2(kA, kA), kB2 = make_float2(kB, kB), kC2 = make_float2(kC, kC);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t rr_[32];
        tmem_ld32(d_col[hh], rr_);
        wait_ld();
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + 384 + ch0[hh]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float2 y[4];
          add_bias8(rr_ + 8 * jj, b4[2 * jj], b4[2 * jj + 1], y);
          uint32_t o4[4];
#pragma unroll
          for (int w2 = 0; w2 < 4; ++w2) {
            const float2 gp = g[16 * hh + 4 * jj + w2];
            float2 ov;
            if constexpr (F2) {
              ov = __ffma2_rn(kA2, gp, __ffma2_rn(kB2, y[w2], kC2));
            } else {
              ov = make_float2(fmaf(kA, gp.x, fmaf(kB, y[w2].x, kC)), fmaf(kA, gp.y, fmaf(kB, y[w2].y, kC)));
            }
            o4[w2] = pack_bf16(ov.x, ov.y);
          }
          *reinterpret_cast<uint4*>(s_gy + tile_off(r, (ch0[hh] >> 3) + jj)) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
    }
    sync_all();
    mark(7);
    if (warp == 0) {
      if (elect_one()) {
        issue_gemm(aW[0], aW[2], true);              // D = gy W4
        issue_wgrad(tmem_base + 384, aW[0], aT[2]);  // dW4 += gy^T a2   (under the epilogue of D)
      }
      __syncwarp();
    }
    colsum(s_gy, acc_b[3]);
    mark(8);
    grad_epilogue(s_T[2]);  // g2 -> T2 in place of a2; waits for the weight-gradient MMA before its first store
    __syncthreads();        // every warp is done reading gy through the generic proxy, and the MMAs reading it are done
    if (tid == 0) {         // gy is dead: bring W2 back into its slot
      mbar_expect_tx(bar_r, kWBlk);
      bulk_g2s(aW[0], p.wpack, kWBlk, bar_r);
    }
    sync_all();
    mark(9);
    if (warp == 0) {
      if (elect_one()) {
        issue_gemm(aT[2], aW[1], true);              // D = g2 W3
        issue_wgrad(tmem_base + 256, aT[2], aT[1]);  // dW3 += g2^T a1
      }
      __syncwarp();
    }
    colsum(s_T[2], acc_b[2]);
    mark(10);
    grad_epilogue(s_T[1]);  // g1 -> T1
    sync_all();
    mark(11);
    if (warp == 0) {
      mbar_wait(bar_r, phase_r);  // W2 is back
      if (elect_one()) {
        issue_gemm(aT[1], aW[0], true);              // D = g1 W2
        issue_wgrad(tmem_base + 128, aT[1], aT[0]);  // dW2 += g1^T a0
      }
      __syncwarp();
    }
    wacc = 1;
    phase_r ^= 1;
    colsum(s_T[1], acc_b[1]);
    mark(12);
    // ---- g0 = D . [a0 > 0] (mask re-derived from the a0 tile) -> fp32 staging over T1|T2, 16-byte
    //      chunks XOR-swizzled by row so that both the row-thread writes and the row-cooperative reads
    //      below are bank-conflict free
    __syncthreads();  // every thread is done with the column sums over T1 (g1) before it is overwritten
    {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        wait_half(hh);
        uint32_t rr_[32];
        tmem_ld32(d_col[hh], rr_);
        wait_ld();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint4 a8 = *reinterpret_cast<const uint4*>(s_T[0] + tile_off(r, (ch0[hh] >> 3) + jj));
          const uint32_t aw[4] = {a8.x, a8.y, a8.z, a8.w};
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t hw = (aw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
            o8[e] = hw ? __uint_as_float(rr_[8 * jj + e]) : 0.f;
          }
          if (hh == 0 && jj == 0) {  // the weight-gradient MMA still reads g1 (T1), which the staging overwrites
            mbar_wait(bar_g, phase_g);
            phase_g ^= 1;
          }
          const int c4 = (ch0[hh] >> 2) + 2 * jj;  // logical 16-byte chunk of the fp32 row
          *reinterpret_cast<float4*>(s_g0 + r * 128 + (((c4 + 0) ^ (r & 31)) << 2)) = make_float4(o8[0], o8[1], o8[2], o8[3]);
          *reinterpret_cast<float4*>(s_g0 + r * 128 + (((c4 + 1) ^ (r & 31)) << 2)) = make_float4(o8[4], o8[5], o8[6], o8[7]);
        }
      }
    }
    __syncthreads();
    mark(13);
    {
      // row-cooperative pass (warp w: rows 16w..16w+15, lane l: channels 4l..4l+3): the gradient of the
      // sender projection goes out as one coalesced 512 B red.add per edge row; the gradient of the
      // receiver projection is reduced over runs of equal dst first (the rows are dst-sorted) so that one
      // red.add per (run, channel) reaches L2; bias / fiber-weight column sums stay in registers
      float4 run = make_float4(0.f, 0.f, 0.f, 0.f);
      int cur = s_ij[warp * 16].y;
      auto scatter_row = [&](int rr) {
        const float4 gv = *reinterpret_cast<const float4*>(s_g0 + rr * 128 + ((lane ^ (rr & 31)) << 2));
        const int2 ij = s_ij[rr];
        const float4 f = s_fib[rr];
        if (DET) {
          // deterministic variant: the row leaves as a row; k_edge_grad_segsum forms gPs | gPd in CSR order
          const long long grow = (long long)tile * 128 + rr;
          if (grow < p.rows) st4(p.g0_rows + (size_t)grow * 128 + 4 * lane, gv);
        } else {
          if (ij.x >= 0) red_add_v4(p.gPsPd + (size_t)ij.x * 256 + 4 * lane, gv.x, gv.y, gv.z, gv.w);
          if (ij.y != cur) {
            if (cur >= 0) red_add_v4(p.gPsPd + (size_t)cur * 256 + 128 + 4 * lane, run.x, run.y, run.z, run.w);
            cur = ij.y;
            run = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          run.x += gv.x; run.y += gv.y; run.z += gv.z; run.w += gv.w;
        }
        acc_b0.x += gv.x; acc_b0.y += gv.y; acc_b0.z += gv.z; acc_b0.w += gv.w;
        acc_fl[0].x += gv.x * f.x; acc_fl[0].y += gv.y * f.x; acc_fl[0].z += gv.z * f.x; acc_fl[0].w += gv.w * f.x;
        acc_fl[1].x += gv.x * f.y; acc_fl[1].y += gv.y * f.y; acc_fl[1].z += gv.z * f.y; acc_fl[1].w += gv.w * f.y;
        acc_fl[2].x += gv.x * f.z; acc_fl[2].y += gv.y * f.z; acc_fl[2].z += gv.z * f.z; acc_fl[2].w += gv.w * f.z;
        acc_fl[3].x += gv.x * f.w; acc_fl[3].y += gv.y * f.w; acc_fl[3].z += gv.z * f.w; acc_fl[3].w += gv.w * f.w;
      };
#pragma unroll 4
      for (int rr = warp * 16; rr < warp * 16 + 16; ++rr) scatter_row(rr);
      if (!DET && cur >= 0) red_add_v4(p.gPsPd + (size_t)cur * 256 + 128 + 4 * lane, run.x, run.y, run.z, run.w);
    }
    __syncthreads();  // the staging tiles / row metadata are rewritten by the next tile
    mark(14);
  }

  if (PROF && tid == 0) {
    for (int k = 0; k < 16; ++k) atomicAdd(p.prof + k, s_prof[k]);
  }
  // ---- flush: weight-gradient accumulators (TMEM) and the per-thread bias / fiber partial sums
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (DET) {
    // per-CTA / per-warp partial sums, added up in a fixed order by det_reduce (layout: kDetEdgeBwdStride)
    float* part = p.part + (size_t)blockIdx.x * kDetEdgeBwdStride;
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
      float v[64];
      load_d64(tmem_base + 128 * (l + 1) + lane_off + 64 * h, v);
      float* dst = part + l * 16384 + r * 128 + 64 * h;
#pragma unroll
      for (int q4 = 0; q4 < 16; ++q4) st4(dst + q4 * 4, make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]));
    }
    float* pb = part + 3 * 16384;
#pragma unroll
    for (int l = 1; l < 4; ++l) st4(pb + ((l - 1) * 8 + warp) * 128 + 4 * lane, acc_b[l]);
    st4(pb + 3 * 8 * 128 + warp * 128 + 4 * lane, acc_b0);
    float* pf = pb + 4 * 8 * 128 + (warp * 128 + 4 * lane) * 4;  // [warp][channel][fiber component]
    st4(pf + 0, make_float4(acc_fl[0].x, acc_fl[1].x, acc_fl[2].x, acc_fl[3].x));
    st4(pf + 4, make_float4(acc_fl[0].y, acc_fl[1].y, acc_fl[2].y, acc_fl[3].y));
    st4(pf + 8, make_float4(acc_fl[0].z, acc_fl[1].z, acc_fl[2].z, acc_fl[3].z));
    st4(pf + 12, make_float4(acc_fl[0].w, acc_fl[1].w, acc_fl[2].w, acc_fl[3].w));
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    return;
  }
  {
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
      float v[64];
      load_d64(tmem_base + 128 * (l + 1) + lane_off + 64 * h, v);
      float* dst = p.gW[l] + (size_t)r * 128 + 64 * h;  // TMEM lane = output channel
#pragma unroll
      for (int q4 = 0; q4 < 16; ++q4) red_add_v4(dst + q4 * 4, v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
    }
  }
#pragma unroll
  for (int l = 1; l < 4; ++l) {
    atomicAdd(p.gb[l] + 4 * lane + 0, acc_b[l].x);
    atomicAdd(p.gb[l] + 4 * lane + 1, acc_b[l].y);
    atomicAdd(p.gb[l] + 4 * lane + 2, acc_b[l].z);
    atomicAdd(p.gb[l] + 4 * lane + 3, acc_b[l].w);
  }
  {
    const int ldw1 = 2 * kD + p.P + 1;
    const float b0[4] = {acc_b0.x, acc_b0.y, acc_b0.z, acc_b0.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      atomicAdd(p.gb[0] + 4 * lane + c, b0[c]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float v = c == 0 ? acc_fl[k].x : (c == 1 ? acc_fl[k].y : (c == 2 ? acc_fl[k].z : acc_fl[k].w));
        if (k <= p.P) atomicAdd(p.gW1 + (size_t)(4 * lane + c) * ldw1 + k, v);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// alignment slack + 3 weight slots + 3 tiles + s_bias[512] + s_F[128] + s_fib[2][128] + s_x[2][128] + s_ij[2][128] + 3 barriers + TMEM address
size_t edge_chain_bwd_smem() { return 1024 + 6 * kWBlk + 512 * 4 + 128 * 16 * 3 + 256 * 16 + 256 * 8 + 5 * 8 + 16 + 128; }

// Fused bf16 backward of the edge stage.  gPsPd must be zero-filled; gW/gb/gW1 are accumulated into.
int edge_chain_backward(const bsms_level_plan* pl, const bsms_gmp_weights* w, const bsms_gmp_grads* gr, const float* PsPd,
                        const float* pos, int pos_batched, int B, int P, uint8_t* wpack, const float* g_aggr, int ld_g,
                        float* gPsPd, cudaStream_t st, bool prepacked, float* g0_rows, float* part) {
  const long long rows = (long long)B * pl->n_edges;
  if (rows == 0) return BSMS_OK;
  PackList pk;
  pk.n = 3;
  for (int l = 0; l < 3; ++l) {
    pk.w[l] = w->w_edge[l + 1];
    pk.ld[l] = kD;
  }
  EdgeBwdParams p;
  p.PsPd = PsPd;
  p.pos = pos;
  p.pos_batched = pos_batched;
  p.P = P;
  p.src_d = pl->src_d;
  p.dst_d = pl->dst_d;
  p.W1 = w->w_edge[0];
  for (int l = 0; l < 4; ++l) {
    p.b[l] = w->b_edge[l];
    p.gb[l] = gr->b_edge[l];
  }
  for (int l = 0; l < 3; ++l) p.gW[l] = gr->w_edge[l + 1];
  p.gW1 = gr->w_edge[0];
  p.wpack = wpack;
  p.g_aggr = g_aggr;
  p.ld_g = ld_g;
  p.gPsPd = gPsPd;
  p.g0_rows = g0_rows;
  p.part = part;
  p.B = B;
  p.N = pl->n_nodes;
  p.E = pl->n_edges;
  p.rows = rows;
  p.ntiles = ceil_div(rows, 128);
  int dev = 0, sms = 148;
  BSMS_CUDA(cudaGetDevice(&dev));
  BSMS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (!prepacked) {
    ProfScope ps_(PK_OTHER, st);
    k_pack_weights<1><<<3, 256, 0, st>>>(pk, wpack);
    BSMS_LAUNCHED();
  }
  static const bool phase_prof = getenv("BSMS_PHASE_PROF") != nullptr;
  static unsigned long long* d_prof = nullptr;
  p.prof = nullptr;
  if (phase_prof) {
    if (!d_prof) BSMS_CUDA(cudaMalloc(&d_prof, 16 * sizeof(unsigned long long)));
    BSMS_CUDA(cudaMemsetAsync(d_prof, 0, 16 * sizeof(unsigned long long), st));
    p.prof = d_prof;
  }
  const size_t smem = edge_chain_bwd_smem();
  // BSMS_BWD_F2=0 (development switch) turns the packed fp32x2 epilogue arithmetic off
  static const bool f2 = !(getenv("BSMS_BWD_F2") && atoi(getenv("BSMS_BWD_F2")) == 0);
  if ((g0_rows == nullptr) != (part == nullptr)) {
    set_error("edge_chain_backward: the deterministic variant needs both the row buffer and the partial-sum block");
    return BSMS_EINVAL;
  }
  void (*kern)(const EdgeBwdParams) =
      g0_rows ? k_edge_chain_bwd<false, true, true>
              : (f2 ? (phase_prof ? k_edge_chain_bwd<true, true, false> : k_edge_chain_bwd<false, true, false>)
                    : (phase_prof ? k_edge_chain_bwd<true, false, false> : k_edge_chain_bwd<false, false, false>));
  BSMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps_(PK_EDGE_CHAIN_BWD, st);
  const int grid = std::min(sms, p.ntiles);
  kern<<<grid, 256, smem, st>>>(p);
  BSMS_LAUNCHED();
  if (part) {
    // ordered reduction of the per-CTA sums into the gradients (layout of the flush above)
    DetSeg segs[8];
    for (int l = 0; l < 3; ++l) segs[l] = DetSeg{p.gW[l], l * 16384, 0, grid, 1, 128, 128, 128, 128};
    for (int l = 1; l < 4; ++l) segs[2 + l] = DetSeg{p.gb[l], 3 * 16384 + (l - 1) * 8 * 128, 0, grid, 8, 1, 128, 128, 128};
    segs[6] = DetSeg{p.gb[0], 3 * 16384 + 3 * 8 * 128, 0, grid, 8, 1, 128, 128, 128};
    segs[7] = DetSeg{p.gW1, 3 * 16384 + 4 * 8 * 128, 0, grid, 8, 128, P + 1, 4, 2 * kD + P + 1};
    return det_reduce(part, kDetEdgeBwdStride, segs, 8, st);
  }
  if (phase_prof) {  // debug aid: per-phase cycles per tile (thread 0 of every CTA), printed per launch
    unsigned long long h[16];
    BSMS_CUDA(cudaMemcpyAsync(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost, st));
    BSMS_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[bwd phases] tiles %d:", p.ntiles);
    for (int k = 0; k < 15; ++k) fprintf(stderr, " %llu", h[k] / (unsigned long long)p.ntiles);
    fprintf(stderr, "\n");
  }
  return BSMS_OK;
}

}  // namespace bsms
