// Fused backward of the node MLP's layers 1..3 on tcgen05 (bf16 operands, fp32 accumulate): the
// LayerNorm backward of the block output, three data-gradient GEMMs with their ReLU masks and three
// weight-gradient GEMMs in ONE persistent kernel.  Nothing between the upstream gradient and the
// gradient of the first layer's activation touches HBM (reference: src/ops/basic.py:6-23,97-98; the
// reference's autograd materialises every intermediate).
//
// Per 128-node-row tile (256 threads; row-cooperative global I/O: one warp per 512 B row):
//   (N2, N3 arrive either as fp32 rows, converted on the way in, or — written by k_node_chain_fwd — as bf16
//    operand-tile images that cp.async.bulk drops straight into T2 / T3)
//   G1 = LN'(Yn) g_out                     rows of Yn, g_out  -> TG (bf16 tile)      ; N3 rows -> T3
//   dV4 += G1^T N3 ; G2 = (G1 V4) . [N3>0]  UMMA wgrad(TG,T3), dgrad D = TG x V4(MN)  -> T3 (in place) ; N2 rows -> T2
//   dV3 += G2^T N2 ; G3 = (G2 V3) . [N2>0]  UMMA wgrad(T3,T2), dgrad D = T3 x V3(MN)  -> T2 (in place) ; N1 rows -> TG
//   dV2 += G3^T N1 ; G4 = (G3 V2) . [N1>0]  UMMA wgrad(T2,TG), dgrad D = T2 x V2(MN)  -> fp32 staging over T3|T2 -> G4 rows
// The loads of the next operand tile are issued in the shadow of the current MMA batch.  The three
// weight-gradient accumulators stay in TMEM for the whole persistent loop (as in edge_chain_bwd.cu);
// bias gradients are column sums of the staged gradient tiles.
#include <stdio.h>
#include <stdlib.h>

#include "chain.cuh"

namespace bsms {

struct NodeBwdParams {
  const float* Yn;     // [rows,128] pre-LayerNorm output of the node MLP (kept from forward)
  const float* g_out;  // [rows,128] gradient of the block output
  const float* N[3];   // N1, N2, N3 [rows,128] (kept from forward)
  int img;             // 1: N[1], N[2] are bf16 operand-tile images (ntiles x 32 KB, written by k_node_chain_fwd)
                       //    and are brought in by cp.async.bulk straight into operand position
  const uint8_t* wpack;  // packed bf16 blocks V2, V3, V4 (contiguous)
  float* G4;           // [rows,128] gradient of the first layer's activation (after its ReLU mask)
  float* gW[3];        // V2, V3, V4 gradients [128,128] (accumulated)
  float* gb[3];        // c2, c3, c4 gradients (accumulated)
  float* part;         // deterministic option: per-CTA sums go to part[blockIdx.x][kDetNodeBwdStride] instead
  int l2_prefetch;     // request the next tile's rows / images into L2 under the first MMA batch of the current tile
  unsigned long long* prof;  // optional [16] per-phase cycle sums of thread 0 of every CTA (BSMS_PHASE_PROF=1)
  long long rows;
  int ntiles;
};

__device__ __forceinline__ uint32_t nt_off(int r, int chunk) {  // 16-byte chunk `chunk` (0..15) of tile row r
  return (uint32_t)((chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(256, 1) k_node_chain_bwd(const NodeBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t IDESC_KM = make_idesc(1, 128, 128, 0, 1);  // A K-major, B MN-major  (dgrad: B = W^T)
  constexpr uint32_t IDESC_MM = make_idesc(1, 128, 128, 1, 1);  // A, B MN-major          (wgrad)
  const uint32_t s0 = smem_u32(smem_raw);
  const uint32_t sbase = (s0 + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sbase - s0);
  // slots: [V2][V3][V4][TG][T3][T2]
  uint8_t* s_TG = sp + 3 * kWBlk;
  uint8_t* s_T3 = sp + 4 * kWBlk;
  uint8_t* s_T2 = sp + 5 * kWBlk;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(sp + 6 * kWBlk);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 3);
  const uint32_t aV[3] = {sbase, sbase + kWBlk, sbase + 2 * kWBlk};
  const uint32_t aTG = sbase + 3 * kWBlk, aT3 = sbase + 4 * kWBlk, aT2 = sbase + 5 * kWBlk;
  float* s_stage = reinterpret_cast<float*>(s_T3);  // fp32 [128][128] over T3|T2

  const int tid = threadIdx.x, warp = (int)uniform(threadIdx.x >> 5), lane = tid & 31;
  const int q = warp & 3, h = warp >> 2, r = q * 32 + lane;
  const uint32_t bar_w = smem_u32(&s_bar[0]), bar_m = smem_u32(&s_bar[1]), bar_l = smem_u32(&s_bar[2]);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    mbar_init(bar_l, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = uniform(*s_tmem);
  if (tid == 0) {
    mbar_expect_tx(bar_w, 3 * kWBlk);
    for (int blk = 0; blk < 3; ++blk) bulk_g2s(aV[blk], p.wpack + (size_t)blk * kWBlk, kWBlk, bar_w);
    mbar_wait(bar_w, 0);
  }
  __syncwarp();  // lane 0 rejoins its warp here: a warp that stays split runs every collective on its slow path
  const uint32_t d_tmem = tmem_base;  // D: cols [0,128); dV2/dV3/dV4: cols [128,256), [256,384), [384,512)
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  const uint32_t d_mine = d_tmem + lane_off + 64 * h;
  uint32_t phase = 0, phase_l = 0, wacc = 0;
  float4 acc_b[3];  // bias-gradient partial sums of this lane's 4 channels (c2, c3, c4)
#pragma unroll
  for (int l = 0; l < 3; ++l) acc_b[l] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto sync_all = [&]() {
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  };
  auto wait_mma = [&]() {
    mbar_wait(bar_m, phase);
    phase ^= 1;
    fence_after_sync();
  };
  // dW[out][in] += G^T A (both MN-major views, K = 128 tile rows) followed by D = G(K-major) x W(MN-major = W^T)
  auto issue_pair = [&](uint32_t dw_tmem, uint32_t g_tile, uint32_t a_tile, uint32_t w_blk) {
    if (warp == 0) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = smem_desc_sw128(g_tile + ks * 2048, 16384, 1024);
          const uint64_t bd = smem_desc_sw128(a_tile + ks * 2048, 16384, 1024);
          mma_ss(dw_tmem, ad, bd, IDESC_MM, (wacc | ks) != 0);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = smem_desc_sw128(g_tile + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
          const uint64_t bd = smem_desc_sw128(w_blk + ks * 2048, 16384, 1024);
          mma_ss(d_tmem, ad, bd, IDESC_KM, ks > 0);
        }
        mma_commit(bar_m);
      }
      __syncwarp();
    }
  };
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
  // rows 16w..16w+15 of X -> bf16 tile (all 16 row requests of the warp in flight at once)
  auto load_tile = [&](const float* X, long long row0, uint8_t* tile) {
    float4 v[16];
    coop_rows_load<16>(X, kD, row0, p.rows, warp * 16, lane, v);
    coop_rows_store<16>(tile, warp * 16, lane, v);
  };
  // gradient epilogue: D . [act > 0] -> the activation tile itself (in place, bf16)
  auto grad_epilogue = [&](uint8_t* tile) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t rr_[32];
      tmem_ld32(d_mine + 32 * hh, rr_);
      wait_ld();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4* cp = reinterpret_cast<uint4*>(tile + nt_off(r, 8 * h + 4 * hh + jj));
        const uint4 a8 = *cp;
        const uint32_t aw[4] = {a8.x, a8.y, a8.z, a8.w};
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t hw = (aw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
          o8[e] = hw ? __uint_as_float(rr_[8 * jj + e]) : 0.f;
        }
        uint4 u;
        u.x = pack_bf16(o8[0], o8[1]); u.y = pack_bf16(o8[2], o8[3]);
        u.z = pack_bf16(o8[4], o8[5]); u.w = pack_bf16(o8[6], o8[7]);
        *cp = u;
      }
    }
  };

  long long tprev = p.prof ? clock64() : 0;
  unsigned long long cyc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  auto mark = [&](int k) {
    if (p.prof && tid == 0) {
      const long long t = clock64();
      cyc[k] += (unsigned long long)(t - tprev);
      tprev = t;
    }
  };
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * 128;
    if (p.img && tid == 0) {
      // N3 -> T3 and N2 -> T2 as ready-made operand tiles (both slots were the previous tile's staging: free)
      mbar_expect_tx(bar_l, 2 * kWBlk);
      bulk_g2s(aT3, reinterpret_cast<const uint8_t*>(p.N[2]) + (size_t)tile * kWBlk, kWBlk, bar_l);
      bulk_g2s(aT2, reinterpret_cast<const uint8_t*>(p.N[1]) + (size_t)tile * kWBlk, kWBlk, bar_l);
    }
    __syncwarp();
    // ---- G1 = LayerNorm backward of g_out through Yn (warp per row), N3 -> T3
    {
      const uint32_t col_off = (uint32_t)((lane >> 4) * 16384 + (lane & 1) * 8);
      const int chunk7 = (lane >> 1) & 7;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float4 y[8], g[8];
        coop_rows_load<8>(p.Yn, kD, row0, p.rows, warp * 16 + 8 * half, lane, y);
        coop_rows_load<8>(p.g_out, kD, row0, p.rows, warp * 16 + 8 * half, lane, g);
        // One pass over FOUR rows at a time: the sums of d = y - shift, d^2, g and g d of the four rows travel through
        // one butterfly — 16 independent shuffles per step, issued back to back, instead of a dependent chain of
        // 5-step reductions per row (the phase was bound by shuffle latency: ~580 cycles per row); the shift (the row's
        // first element) keeps E[d^2] - E[d]^2 free of cancellation, as in the forward kernels
#pragma unroll
        for (int u0 = 0; u0 < 8; u0 += 4) {
          float sh[4], s1[4], s2[4], s3[4], s4[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) sh[v] = __shfl_sync(0xffffffffu, y[u0 + v].x, 0);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 yy = y[u0 + v], gg = g[u0 + v];
            const float dx = yy.x - sh[v], dy = yy.y - sh[v], dz = yy.z - sh[v], dw = yy.w - sh[v];
            s1[v] = (dx + dy) + (dz + dw);
            s2[v] = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
            s3[v] = (gg.x + gg.y) + (gg.z + gg.w);
            s4[v] = fmaf(gg.x, dx, fmaf(gg.y, dy, fmaf(gg.z, dz, gg.w * dw)));
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            float t1[4], t2[4], t3[4], t4[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              t1[v] = __shfl_xor_sync(0xffffffffu, s1[v], o);
              t2[v] = __shfl_xor_sync(0xffffffffu, s2[v], o);
              t3[v] = __shfl_xor_sync(0xffffffffu, s3[v], o);
              t4[v] = __shfl_xor_sync(0xffffffffu, s4[v], o);
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              s1[v] += t1[v]; s2[v] += t2[v]; s3[v] += t3[v]; s4[v] += t4[v];
            }
          }
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int rr = warp * 16 + 8 * half + u0 + v;
            const float4 yy = y[u0 + v], gg = g[u0 + v];
            const float md = s1[v] * (1.f / 128.f);
            const float var = fmaxf(s2[v] * (1.f / 128.f) - md * md, 0.f);
            const float rstd = 1.f / sqrtf(var + 1e-5f);
            const float c1 = s3[v] * (1.f / 128.f);
            const float c2 = rstd * (s4[v] * (1.f / 128.f) - md * c1);  // mean over the row of g h, h = (d - md) rstd
            const float m = sh[v] + md;
            const float hx = (yy.x - m) * rstd, hy = (yy.y - m) * rstd, hz = (yy.z - m) * rstd, hw_ = (yy.w - m) * rstd;
            uint2 pk;  // rows past the end load zeros: y = g = 0 gives G1 = 0
            pk.x = pack_bf16(rstd * (gg.x - c1 - hx * c2), rstd * (gg.y - c1 - hy * c2));
            pk.y = pack_bf16(rstd * (gg.z - c1 - hz * c2), rstd * (gg.w - c1 - hw_ * c2));
            *reinterpret_cast<uint2*>(s_TG + col_off + rr * 128 + ((chunk7 ^ (rr & 7)) << 4)) = pk;
          }
        }

      }
    }
    if (p.prof && tid == 0 && tile == (int)blockIdx.x) cyc[9] = (unsigned long long)(clock64() - tprev);
    mark(0);
    if (p.img) {
      mbar_wait(bar_l, phase_l);  // the two image tiles have landed
      phase_l ^= 1;
    } else {
      load_tile(p.N[2], row0, s_T3);
    }
    sync_all();
    mark(1);
    issue_pair(tmem_base + 384, aTG, aT3, aV[2]);  // dV4 += G1^T N3 ; D = G1 V4
    if (p.l2_prefetch && tile + (int)gridDim.x < p.ntiles) {
      // the next tile's rows are requested into L2 now (fire and forget): its three dependent load phases then see
      // L2 latency instead of DRAM latency.  128 rows x 4 lines x 3 tensors (+ 2 x 256 lines of images) over 256 threads
      const long long nrow0 = (long long)(tile + gridDim.x) * 128;
      const long long lim = p.rows * kD;  // floats
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const long long off = nrow0 * kD + (long long)(tid + 256 * k) * 32;  // 32 floats = one 128-byte line
        if (off < lim) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.Yn + off));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.g_out + off));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.N[0] + off));
        }
      }
      if (p.img) {
        const size_t ioff = (size_t)(tile + gridDim.x) * kWBlk + (size_t)tid * 128;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(p.N[2]) + ioff));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(p.N[1]) + ioff));
      }
    }
    colsum(s_TG, acc_b[2]);
    if (!p.img) load_tile(p.N[1], row0, s_T2);     // N2 -> T2 in the shadow of the MMAs
    wait_mma();
    mark(2);
    grad_epilogue(s_T3);                           // G2 -> T3
    sync_all();
    mark(3);
    issue_pair(tmem_base + 256, aT3, aT2, aV[1]);  // dV3 += G2^T N2 ; D = G2 V3
    colsum(s_T3, acc_b[1]);
    load_tile(p.N[0], row0, s_TG);                 // N1 -> TG (G1 is dead: its MMAs completed)
    wait_mma();
    mark(4);
    grad_epilogue(s_T2);                           // G3 -> T2
    sync_all();
    mark(5);
    issue_pair(tmem_base + 128, aT2, aTG, aV[0]);  // dV2 += G3^T N1 ; D = G3 V2
    wacc = 1;
    colsum(s_T2, acc_b[0]);
    wait_mma();
    __syncthreads();  // every warp is done with the column sums over T2 before the staging overwrites it
    mark(6);
    // ---- G4 = D . [N1 > 0] -> fp32 staging over T3|T2 (16-byte chunks XOR-swizzled by row)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t rr_[32];
      tmem_ld32(d_mine + 32 * hh, rr_);
      wait_ld();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint4 a8 = *reinterpret_cast<const uint4*>(s_TG + nt_off(r, 8 * h + 4 * hh + jj));
        const uint32_t aw[4] = {a8.x, a8.y, a8.z, a8.w};
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t hw = (aw[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
          o8[e] = hw ? __uint_as_float(rr_[8 * jj + e]) : 0.f;
        }
        const int c4 = 16 * h + 8 * hh + 2 * jj;
        *reinterpret_cast<float4*>(s_stage + r * 128 + (((c4 + 0) ^ (r & 31)) << 2)) = make_float4(o8[0], o8[1], o8[2], o8[3]);
        *reinterpret_cast<float4*>(s_stage + r * 128 + (((c4 + 1) ^ (r & 31)) << 2)) = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
    }
    __syncthreads();
    mark(7);
#pragma unroll 4
    for (int rr = warp * 16; rr < warp * 16 + 16; ++rr) {
      const long long row = row0 + rr;
      if (row < p.rows)
        st4(p.G4 + row * kD + 4 * lane, *reinterpret_cast<const float4*>(s_stage + rr * 128 + ((lane ^ (rr & 31)) << 2)));
    }
    sync_all();  // the tiles are rewritten by the next tile (by cp.async.bulk in image mode: proxy fence included)
    mark(8);
  }
  if (p.prof && tid == 0)
    for (int k = 0; k < 10; ++k) atomicAdd(p.prof + k, cyc[k]);

  // ---- flush: weight-gradient accumulators (TMEM) and the per-lane bias partial sums
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (p.part) {
    // deterministic option: plain stores of this CTA's sums (zeros when it had no tile); det_reduce adds them in order
    float* part = p.part + (size_t)blockIdx.x * kDetNodeBwdStride;
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t rr_[32];
        if (wacc) {
          tmem_ld32(tmem_base + 128 * (l + 1) + lane_off + 64 * h + 32 * hh, rr_);
          wait_ld();
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) rr_[t] = 0u;
        }
        float* dst = part + l * 16384 + r * 128 + 64 * h + 32 * hh;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4)
          st4(dst + q4 * 4, make_float4(__uint_as_float(rr_[q4 * 4]), __uint_as_float(rr_[q4 * 4 + 1]),
                                        __uint_as_float(rr_[q4 * 4 + 2]), __uint_as_float(rr_[q4 * 4 + 3])));
      }
      st4(part + 3 * 16384 + (l * 8 + warp) * 128 + 4 * lane, acc_b[l]);
    }
  } else if (wacc) {
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t rr_[32];
        tmem_ld32(tmem_base + 128 * (l + 1) + lane_off + 64 * h + 32 * hh, rr_);
        wait_ld();
        float* dst = p.gW[l] + (size_t)r * 128 + 64 * h + 32 * hh;  // TMEM lane = output channel
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4)
          red_add_v4(dst + q4 * 4, __uint_as_float(rr_[q4 * 4]), __uint_as_float(rr_[q4 * 4 + 1]),
                     __uint_as_float(rr_[q4 * 4 + 2]), __uint_as_float(rr_[q4 * 4 + 3]));
      }
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      atomicAdd(p.gb[l] + 4 * lane + 0, acc_b[l].x);
      atomicAdd(p.gb[l] + 4 * lane + 1, acc_b[l].y);
      atomicAdd(p.gb[l] + 4 * lane + 2, acc_b[l].z);
      atomicAdd(p.gb[l] + 4 * lane + 3, acc_b[l].w);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// G4 = gradient of the first node layer's activation; accumulates dV2..dV4 and dc2..dc4.
int node_chain_backward(const float* Yn, const float* g_out, const float* N1, const float* N2, const float* N3, int img,
                        const uint8_t* wpack_v2, float* G4, float* const* gW, float* const* gb, long long rows,
                        cudaStream_t st, float* part) {
  if (rows == 0) return BSMS_OK;
  NodeBwdParams p;
  p.Yn = Yn;
  p.g_out = g_out;
  p.N[0] = N1;
  p.N[1] = N2;
  p.N[2] = N3;
  p.img = img;
  p.wpack = wpack_v2;
  p.G4 = G4;
  for (int l = 0; l < 3; ++l) {
    p.gW[l] = gW[l];
    p.gb[l] = gb[l];
  }
  p.part = part;
  static const int l2pf = getenv("BSMS_NODE_PF") ? atoi(getenv("BSMS_NODE_PF")) : 0;  // experiment switch
  p.l2_prefetch = l2pf;
  p.rows = rows;
  p.ntiles = ceil_div(rows, 128);
  int dev = 0, sms = 148;
  BSMS_CUDA(cudaGetDevice(&dev));
  BSMS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = 1024 + 6 * kWBlk + 3 * 8 + 16;
  BSMS_CUDA(cudaFuncSetAttribute(k_node_chain_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps_(PK_DGRAD, st);
  const int grid = std::min(sms, p.ntiles);
  static const bool phase_prof = getenv("BSMS_PHASE_PROF") != nullptr;
  static unsigned long long* d_prof = nullptr;
  p.prof = nullptr;
  if (phase_prof) {
    if (!d_prof) BSMS_CUDA(cudaMalloc(&d_prof, 16 * sizeof(unsigned long long)));
    BSMS_CUDA(cudaMemsetAsync(d_prof, 0, 16 * sizeof(unsigned long long), st));
    p.prof = d_prof;
  }
  k_node_chain_bwd<<<grid, 256, smem, st>>>(p);
  BSMS_LAUNCHED();
  if (phase_prof) {  // debug aid: per-phase cycles per tile: 0 LN backward (loads + math), 1 wait N3/N2 images, 2 MMA pair 1 +
                     // column sums, 3 epilogue 1, 4 MMA pair 2 + N1 rows, 5 epilogue 2, 6 MMA pair 3, 7 staging, 8 row stores
    unsigned long long h[16];
    BSMS_CUDA(cudaMemcpyAsync(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost, st));
    BSMS_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[node bwd phases] tiles %d:", p.ntiles);
    for (int k = 0; k < 9; ++k) fprintf(stderr, " %llu", h[k] / (unsigned long long)p.ntiles);
    fprintf(stderr, " | phase 0 of a CTA's first tile: %llu", h[9] / (unsigned long long)grid);
    fprintf(stderr, "\n");
  }
  if (part) {
    DetSeg segs[6];
    for (int l = 0; l < 3; ++l) {
      segs[l] = DetSeg{gW[l], l * 16384, 0, grid, 1, 128, 128, 128, 128};
      segs[3 + l] = DetSeg{gb[l], 3 * 16384 + l * 8 * 128, 0, grid, 8, 1, 128, 128, 128};
    }
    return det_reduce(part, kDetNodeBwdStride, segs, 6, st);
  }
  return BSMS_OK;
}

}  // namespace bsms
