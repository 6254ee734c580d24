// Node-level dense layers on tcgen05: the per-node pre-projection of the first edge layer, the
// node MLP, and their data / weight gradients (reference: src/ops/basic.py:6-23, :97-98).
//
//   k_lin<NSPLIT>  Y[rows, NB*128] = epi( [X0 | X1][rows, KB*128] x W )      TS-form, thread = row = TMEM lane
//                  - forward form : W blocks are [n][k] images read K-major
//                  - dgrad form   : W blocks are [k][n] images read MN-major (= W^T), no transposed copy
//                  - epilogue     : +bias, ReLU, mask by (M > 0) (ReLU backward), += Y
//                  One warpgroup per CTA, 256 TMEM columns per CTA, two CTAs per SM ping-pong.
//   k_wgrad_tc     dW[128,128] += G^T X over ALL rows: both fp32 row tiles are converted to bf16
//                  canonical tiles in shared memory (double buffered) and consumed MN-major; the
//                  accumulator stays in TMEM for the whole persistent loop; bias gradient = column sums.
#include "chain.cuh"

namespace bsms {

struct LinParams {
  const float* X[2];
  int ldx[2];
  int KB, NB;
  const uint8_t* wblk[4];  // packed block of (nb, kb) at index nb*2+kb
  int b_mn;                // 1: blocks are [k][n] images, read MN-major (dgrad form)
  const float* bias;
  int relu;
  const float* mask;
  int ldmask;
  int accum;
  float* Y;
  int ldy;
  long long rows;
  int ntiles;
};

template <int NSPLIT>
__global__ void __launch_bounds__(128, 2) k_lin(const LinParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t FMT = (NSPLIT == 1) ? 1u : 0u;
  constexpr uint32_t BLK = NSPLIT * kWBlk;
  constexpr float OUT_SCALE = (NSPLIT == 1) ? 1.f : 1.f / (kActScale * kWScale);
  const uint32_t idesc = make_idesc(FMT, 128, 128, 0, p.b_mn ? 1u : 0u);
  const uint32_t s0 = smem_u32(smem_raw);
  const uint32_t sbase = (s0 + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sbase - s0);
  const int nblk = p.KB * p.NB;
  float* s_bias = reinterpret_cast<float*>(sp + 2 * BLK);  // [256]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2);
  const int tid = threadIdx.x, warp = (int)uniform(threadIdx.x >> 5);
  const uint32_t bar_w = smem_u32(&s_bar[0]), bar_m = smem_u32(&s_bar[1]);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 256);
  for (int i = tid; i < 128 * p.NB; i += 128) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = uniform(*s_tmem);
  if (tid == 0) {
    mbar_expect_tx(bar_w, nblk * BLK);
    for (int nb = 0; nb < p.NB; ++nb)
      for (int kb = 0; kb < p.KB; ++kb)
        bulk_g2s(sbase + (nb * p.KB + kb) * BLK, p.wblk[nb * 2 + kb], BLK, bar_w);
    mbar_wait(bar_w, 0);
  }
  __syncwarp();  // lane 0 rejoins its warp (a warp left split runs its collectives on the slow path until the next barrier)
  const uint32_t d_tmem = tmem_base;       // 128 columns
  const uint32_t a_tmem = tmem_base + 128;  // 64 (hi) [+ 64 (lo)]
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const long long row = (long long)tile * 128 + tid;
    const bool valid = row < p.rows;
    for (int nb = 0; nb < p.NB; ++nb) {
      for (int kb = 0; kb < p.KB; ++kb) {
        if (nb == 0 || p.KB > 1) {
          const float* xr = p.X[kb] + row * p.ldx[kb];
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            float v[32];
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              float4 a = valid ? ld4(xr + c0 + q4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              v[q4 * 4 + 0] = a.x; v[q4 * 4 + 1] = a.y; v[q4 * 4 + 2] = a.z; v[q4 * 4 + 3] = a.w;
            }
            uint32_t hi[16];
            if (NSPLIT == 1) {
#pragma unroll
              for (int t = 0; t < 16; ++t) hi[t] = pack_bf16(v[2 * t], v[2 * t + 1]);
              tmem_st16(a_tmem + lane_off + (c0 >> 1), hi);
            } else {
              uint32_t lo[16];
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                float s0_ = v[2 * t] * kActScale, s1_ = v[2 * t + 1] * kActScale;
                __half h0 = __float2half_rn(s0_), h1 = __float2half_rn(s1_);
                hi[t] = pack_f16(h0, h1);
                lo[t] = pack_f16(__float2half_rn(s0_ - __half2float(h0)), __float2half_rn(s1_ - __half2float(h1)));
              }
              tmem_st16(a_tmem + lane_off + (c0 >> 1), hi);
              tmem_st16(a_tmem + lane_off + 64 + (c0 >> 1), lo);
            }
          }
          wait_st();
        }
        fence_before_sync();
        __syncthreads();
        if (warp == 0) {  // one elected lane issues; operands are warp-uniform
          fence_after_sync();
          if (elect_one()) {
            const uint32_t wb = sbase + (nb * p.KB + kb) * BLK;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t off = p.b_mn ? ks * 2048 : (ks >> 2) * 16384 + (ks & 3) * 32;
              const uint32_t lbo = p.b_mn ? 16384 : 16;
              const uint64_t bhi = smem_desc_sw128(wb + off, lbo, 1024);
              mma_ts(d_tmem, a_tmem + ks * 8, bhi, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
              if (NSPLIT == 2) {
                const uint64_t blo = smem_desc_sw128(wb + kWBlk + off, lbo, 1024);
                mma_ts(d_tmem, a_tmem + 64 + ks * 8, bhi, idesc, 1);
                mma_ts(d_tmem, a_tmem + ks * 8, blo, idesc, 1);
              }
            }
            mma_commit(bar_m);
          }
          __syncwarp();
        }
        mbar_wait(bar_m, phase);
        phase ^= 1;
        fence_after_sync();
      }
      // ---- epilogue of N block nb
      float* yr = p.Y + row * p.ldy + nb * 128;
      const float* mr = p.mask ? p.mask + row * p.ldmask + nb * 128 : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(d_tmem + lane_off + c0, r);
        wait_ld();
        if (valid) {
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = __uint_as_float(r[q4 * 4 + e]) * OUT_SCALE + s_bias[nb * 128 + c0 + q4 * 4 + e];
              o[e] = p.relu ? fmaxf(x, 0.f) : x;
            }
            if (mr) {
              float4 m = ld4(mr + c0 + q4 * 4);
              o[0] = m.x > 0.f ? o[0] : 0.f; o[1] = m.y > 0.f ? o[1] : 0.f;
              o[2] = m.z > 0.f ? o[2] : 0.f; o[3] = m.w > 0.f ? o[3] : 0.f;
            }
            if (p.accum) {
              float4 y = ld4(yr + c0 + q4 * 4);
              o[0] += y.x; o[1] += y.y; o[2] += y.z; o[3] += y.w;
            }
            st4(yr + c0 + q4 * 4, make_float4(o[0], o[1], o[2], o[3]));
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------
// k_lin2 (bf16): the same contract as k_lin with every global access row-cooperative (one warp per
// 512 B row: 4 L1 wavefronts per request instead of the 32 of a lane-per-row access pattern):
//   load   : warp w brings rows 16w..16w+15 of the tile (lane l = channels 4l..4l+3), converts to bf16 and
//            writes the K-major SWIZZLE_128B A tile in shared memory -> SS-form UMMA; K block 0 of the NEXT
//            tile is prefetched into registers while the current tile's MMAs and epilogue run
//   epilogue: thread (row, column half) reads its TMEM lane, applies bias / ReLU and stages fp32 into a
//            row-swizzled shared-memory tile (aliases the A tiles, dead once the MMAs completed);
//   store  : row-cooperative pass over the staged tile: ReLU-backward mask, addend, coalesced 512 B
//            stores, and optionally the fused LayerNorm + residual of the node MLP's last layer
//            (out = LN(y) + x [+ skip], reference src/ops/basic.py:18,98 and BSMS.py:102).
// 256 threads; 64 KB (A / staging) + 32 KB per weight block: two CTAs per SM for the 1-block layers.
struct Lin2Params {
  const float* X[2];
  int ldx[2];
  int KB, NB;
  const uint8_t* wblk[4];  // packed block of (nb, kb) at index nb*2+kb
  int b_mn;
  const float* bias;  // [NB*128] or null
  int relu;
  const float* mask;  // ReLU backward: keep where mask > 0 (column offset nb*128)
  int ldmask;
  const float* add[2];  // optional addend of output block nb
  int ldadd[2];
  float* Y[2];
  int ldy[2];
  float a_scale;                // SPLIT: power-of-two scale of the A operand (activations: kActScale) ...
  const unsigned* a_amax_dev;   // ... or, when non-null, derived from the operand's max |value| (float bits) in device memory:
                                // gradients, whose magnitude follows the loss — the producer of the tensor recorded it
  unsigned* y_amax_dev;         // when non-null: atomicMax of the float bits of max |stored output| (the next GEMM's a_amax_dev)
  float* ln_out;  // NB == 1: ln_out = LN(y) + res0 (+ res1), rows of 128
  const float* res0;
  const float* res1;
  long long rows;
  int ntiles;
};

// SPLIT = 0: bf16 operands (one MMA per K step).  SPLIT = 1: both operands as a two-way fp16 split of a power-of-two
// scaled value (hi + lo, 22 significant bits), three MMAs per K step (hi.hi + lo.hi + hi.lo), the accumulator is scaled
// back in the epilogue; weight blocks are the [hi | lo] pairs of k_pack_weights<2> (scale kWScale); the A scale is
// kActScale for activations and a device-resident scale for gradients (whose magnitude follows the loss).  KB must be 1
// (a K = 256 layer is two launches, the second adding the first).
// power-of-two scale that puts a tensor's largest magnitude (float bits in device memory) at 512..1024: the hi piece of
// the fp16 split is normal for everything down to max * 2^-24 and nothing overflows (accumulation is fp32 in TMEM)
__device__ __forceinline__ float split_scale(unsigned amax_bits) {
  const float m = __uint_as_float(amax_bits);
  return (m > 0.f && m < 3.0e38f) ? exp2f(floorf(log2f(1024.f / m))) : 1.f;
}
// max over the warp, then one atomicMax of the float bits (non-negative floats order like their bit patterns) — skipped
// when the slot already holds something at least as large
__device__ __forceinline__ void publish_amax(unsigned* slot, float m) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && __float_as_uint(m) > *reinterpret_cast<volatile unsigned*>(slot))
    atomicMax(slot, __float_as_uint(m));
}

template <int SPLIT>
__global__ void __launch_bounds__(256, SPLIT ? 1 : 2) k_lin2(const Lin2Params p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t NSP = SPLIT ? 2 : 1;
  constexpr uint32_t WB = NSP * kWBlk;  // bytes of one packed weight block
  const uint32_t idesc = make_idesc(SPLIT ? 0u : 1u, 128, 128, 0, p.b_mn ? 1u : 0u);
  const float a_scale = SPLIT ? (p.a_amax_dev ? split_scale(*p.a_amax_dev) : p.a_scale) : 1.f;
  float y_amax = 0.f;
  const float out_scale = SPLIT ? 1.f / (a_scale * kWScale) : 1.f;
  const uint32_t s0 = smem_u32(smem_raw);
  const uint32_t sbase = (s0 + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sbase - s0);
  const int nblk = p.KB * p.NB;
  // [A tiles / fp32 staging: 64 KB][weight blocks: nblk x 32 KB][bias 256 floats][barriers]
  uint8_t* s_A = sp;
  float* s_stage = reinterpret_cast<float*>(sp);
  const uint32_t a_addr = sbase, w_addr = sbase + 2 * kWBlk;
  float* s_bias = reinterpret_cast<float*>(sp + 2 * kWBlk + nblk * WB);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2);
  const int tid = threadIdx.x, warp = (int)uniform(threadIdx.x >> 5), lane = tid & 31;
  const uint32_t bar_w = smem_u32(&s_bar[0]), bar_m = smem_u32(&s_bar[1]);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    fence_mbar_init();
  }
  const uint32_t ncols = p.NB == 2 ? 256u : 128u;
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), ncols);
  for (int i = tid; i < 128 * p.NB; i += 256) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = uniform(*s_tmem);
  if (tid == 0) {
    mbar_expect_tx(bar_w, nblk * WB);
    for (int nb = 0; nb < p.NB; ++nb)
      for (int kb = 0; kb < p.KB; ++kb)
        bulk_g2s(w_addr + (nb * p.KB + kb) * WB, p.wblk[nb * 2 + kb], WB, bar_w);
  }
  const int q = warp & 3, h = warp >> 2, r = q * 32 + lane;
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  uint32_t phase = 0;
  bool weights_ready = false;
  float4 pre[16];  // K block 0 of the next tile, rows 16*warp .. 16*warp+15
  if ((int)blockIdx.x < p.ntiles) coop_rows_load<16>(p.X[0], p.ldx[0], (long long)blockIdx.x * 128, p.rows, warp * 16, lane, pre);

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * 128;
    if (SPLIT)
      coop_rows_store_split<16>(s_A, s_A + kWBlk, warp * 16, lane, pre, a_scale);  // A hi | A lo (KB == 1)
    else
      coop_rows_store<16>(s_A, warp * 16, lane, pre);
    if (!SPLIT && p.KB == 2) {
      float4 v1[16];
      coop_rows_load<16>(p.X[1], p.ldx[1], row0, p.rows, warp * 16, lane, v1);
      coop_rows_store<16>(s_A + kWBlk, warp * 16, lane, v1);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0) {  // one elected lane issues; operands are warp-uniform (no per-lane R2UR loop per MMA)
      if (!weights_ready) {
        mbar_wait(bar_w, 0);
        weights_ready = true;
      }
      fence_after_sync();
      // the loops run in warp-uniform control flow (whole warp), only the MMAs sit under the elected lane
      for (int nb = 0; nb < p.NB; ++nb)
        for (int kb = 0; kb < p.KB; ++kb) {
          const uint32_t wb = w_addr + (nb * p.KB + kb) * WB;
          const uint32_t ab = a_addr + kb * kWBlk;
          const uint32_t dt = tmem_base + nb * 128;
          const uint32_t acc0 = kb > 0 ? 1u : 0u;
          if (p.b_mn) {
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint64_t ad = smem_desc_sw128(ab + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
                const uint64_t bd = smem_desc_sw128(wb + ks * 2048, 16384, 1024);
                mma_ss(dt, ad, bd, idesc, ks > 0 ? 1u : acc0);
                if (SPLIT) {
                  mma_ss(dt, smem_desc_sw128(ab + kWBlk + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), bd, idesc, 1u);  // lo . hi
                  mma_ss(dt, ad, smem_desc_sw128(wb + kWBlk + ks * 2048, 16384, 1024), idesc, 1u);                      // hi . lo
                }
              }
            }
          } else {
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint32_t ko = (ks >> 2) * 16384 + (ks & 3) * 32;
                const uint64_t ad = smem_desc_sw128(ab + ko, 16, 1024);
                const uint64_t bd = smem_desc_sw128(wb + ko, 16, 1024);
                mma_ss(dt, ad, bd, idesc, ks > 0 ? 1u : acc0);
                if (SPLIT) {
                  mma_ss(dt, smem_desc_sw128(ab + kWBlk + ko, 16, 1024), bd, idesc, 1u);
                  mma_ss(dt, ad, smem_desc_sw128(wb + kWBlk + ko, 16, 1024), idesc, 1u);
                }
              }
            }
          }
          __syncwarp();
        }
      if (elect_one()) mma_commit(bar_m);
      __syncwarp();
    }
    {
      const int next = tile + gridDim.x;
      if (next < p.ntiles) coop_rows_load<16>(p.X[0], p.ldx[0], (long long)next * 128, p.rows, warp * 16, lane, pre);
    }
    mbar_wait(bar_m, phase);
    phase ^= 1;
    fence_after_sync();
    for (int nb = 0; nb < p.NB; ++nb) {
      // ---- TMEM -> (+bias, ReLU) -> fp32 staging; 16-byte chunks XOR-swizzled by row (conflict free both ways)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t rr_[32];
        tmem_ld32(tmem_base + nb * 128 + lane_off + 64 * h + 32 * hh, rr_);
        wait_ld();
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float acc = SPLIT ? __uint_as_float(rr_[q4 * 4 + e]) * out_scale : __uint_as_float(rr_[q4 * 4 + e]);
            const float x = acc + s_bias[nb * 128 + 64 * h + 32 * hh + q4 * 4 + e];
            o[e] = p.relu ? fmaxf(x, 0.f) : x;
          }
          const int c4 = 16 * h + 8 * hh + q4;
          *reinterpret_cast<float4*>(s_stage + r * 128 + ((c4 ^ (r & 31)) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      __syncthreads();
      // ---- row-cooperative store pass
      const float* mk = p.mask ? p.mask + nb * 128 + 4 * lane : nullptr;
      const float* ad = p.add[nb] ? p.add[nb] + 4 * lane : nullptr;
      float* yo = p.Y[nb] + 4 * lane;
#pragma unroll 4
      for (int rr = warp * 16; rr < warp * 16 + 16; ++rr) {
        const long long row = row0 + rr;
        if (row < p.rows) {
          float4 v = *reinterpret_cast<const float4*>(s_stage + rr * 128 + ((lane ^ (rr & 31)) << 2));
          if (mk) {
            const float4 m = ld4(mk + row * p.ldmask);
            v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
          }
          if (ad) {
            const float4 a = ld4(ad + row * p.ldadd[nb]);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
          }
          st4(yo + row * p.ldy[nb], v);
          if (p.y_amax_dev) y_amax = fmaxf(y_amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
          if (p.ln_out) {
            const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / 128.f);
            const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
            const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / 128.f);
            const float rstd = 1.f / sqrtf(var + 1e-5f);
            float4 o = p.res0 ? ld4(p.res0 + row * 128 + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);  // no residual: plain LayerNorm
            o.x += dx * rstd; o.y += dy * rstd; o.z += dz * rstd; o.w += dw * rstd;
            if (p.res1) {
              const float4 s = ld4(p.res1 + row * 128 + 4 * lane);
              o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
            }
            st4(p.ln_out + row * 128 + 4 * lane, o);
          }
        }
      }
      __syncthreads();  // the staging tile is rewritten by the next block / the next tile's A tiles
    }
  }
  if (p.y_amax_dev) publish_amax(p.y_amax_dev, y_amax);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------
// several independent weight-gradient problems in one launch (CTA c works on problem c / ctas_per_prob)
struct WgradBatch {
  WgradParams prob[6];
  int nprob;
  int ctas_per_prob;
  float* part;  // deterministic option: per-CTA sums go to part[blockIdx.x][kDetWgradStride] instead of atomics
};

__device__ __forceinline__ uint32_t t_off(int r, int chunk) {
  return (uint32_t)((chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(256, 1) k_wgrad_tc(const WgradBatch batch) {
  extern __shared__ uint8_t smem_raw[];
  const int cpp = batch.ctas_per_prob;
  const WgradParams p = batch.prob[blockIdx.x / cpp];
  const int cta_in_prob = blockIdx.x % cpp;
  constexpr uint32_t IDESC_MM = make_idesc(1, 128, 128, 1, 1);
  const uint32_t s0 = smem_u32(smem_raw);
  const uint32_t sbase = (s0 + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sbase - s0);
  // stage s: G tile at (2s) * 32 KB, X tile at (2s+1) * 32 KB
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(sp + 4 * kWBlk);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2);
  const int tid = threadIdx.x, warp = (int)uniform(threadIdx.x >> 5), lane = tid & 31;
  const uint32_t bar0 = smem_u32(&s_bar[0]), bar1 = smem_u32(&s_bar[1]);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 128);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = uniform(*s_tmem);
  const int cc = tid & 127, rh = tid >> 7;
  float acc_b = 0.f;
  uint32_t ph[2] = {0u, 0u};
  int it = 0;
  for (int tile = cta_in_prob; tile < p.ntiles; tile += cpp, ++it) {
    const int s = it & 1;
    // the MMAs that read stage s two iterations ago must be done before it is overwritten
    if (it >= 2) {
      mbar_wait(s ? bar1 : bar0, ph[s]);
      ph[s] ^= 1;
    }
    uint8_t* tg = sp + (2 * s) * kWBlk;
    uint8_t* tx = sp + (2 * s + 1) * kWBlk;
    // row-cooperative loads: warp w brings rows 16w..16w+15 of both tiles as coalesced 512 B rows
    {
      const long long row0 = (long long)tile * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 vg[8], vx[8];
        coop_rows_load<8>(p.G, p.ldg, row0, p.rows, warp * 16 + 8 * half, lane, vg);
        coop_rows_load<8>(p.X, p.ldx, row0, p.rows, warp * 16 + 8 * half, lane, vx);
        coop_rows_store<8>(tg, warp * 16 + 8 * half, lane, vg);
        coop_rows_store<8>(tx, warp * 16 + 8 * half, lane, vx);
      }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      fence_after_sync();
      if (elect_one()) {
        const uint32_t ag = sbase + (2 * s) * kWBlk, ax = sbase + (2 * s + 1) * kWBlk;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = smem_desc_sw128(ag + ks * 2048, 16384, 1024);
          const uint64_t bd = smem_desc_sw128(ax + ks * 2048, 16384, 1024);
          mma_ss(tmem_base, ad, bd, IDESC_MM, (it > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit(s ? bar1 : bar0);
      }
      __syncwarp();
    }
    if (p.db) {  // bias gradient: column sums of the (bf16-rounded) G tile, overlapping the MMAs
      float sacc = 0.f;
#pragma unroll 8
      for (int rr = rh * 64; rr < rh * 64 + 64; ++rr) {
        const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(tg + t_off(rr, cc >> 3) + (cc & 7) * 2);
        sacc += __bfloat162float(*e);
      }
      acc_b += sacc;
    }
  }
  // drain: wait for the last (up to two) commits
  const int n_it = it;
  for (int k = (n_it >= 2 ? n_it - 2 : 0); k < n_it; ++k) {
    const int s = k & 1;
    mbar_wait(s ? bar1 : bar0, ph[s]);
    ph[s] ^= 1;
  }
  fence_after_sync();
  if (n_it > 0) {
    const int q = warp & 3, hh = warp >> 2;  // TMEM lane = output channel n
    const int n = q * 32 + lane;
    uint32_t r0[32], r1[32];
    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + 64 * hh;
    tmem_ld32(ta, r0);
    tmem_ld32(ta + 32, r1);
    wait_ld();
    if (batch.part) {
      float* part = batch.part + (size_t)blockIdx.x * kDetWgradStride;
      float* dst = part + n * 128 + 64 * hh;
#pragma unroll
      for (int t = 0; t < 32; t += 4)
        st4(dst + t, make_float4(__uint_as_float(r0[t]), __uint_as_float(r0[t + 1]), __uint_as_float(r0[t + 2]), __uint_as_float(r0[t + 3])));
#pragma unroll
      for (int t = 0; t < 32; t += 4)
        st4(dst + 32 + t, make_float4(__uint_as_float(r1[t]), __uint_as_float(r1[t + 1]), __uint_as_float(r1[t + 2]), __uint_as_float(r1[t + 3])));
      part[16384 + rh * 128 + cc] = acc_b;
    } else {
      float* dst = p.dW + (size_t)n * p.ldo + 64 * hh;
#pragma unroll
      for (int t = 0; t < 32; ++t) atomicAdd(dst + t, __uint_as_float(r0[t]));
#pragma unroll
      for (int t = 0; t < 32; ++t) atomicAdd(dst + 32 + t, __uint_as_float(r1[t]));
      if (p.db) atomicAdd(p.db + cc, acc_b);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------ host side
static int g_sms = 0;
static int sm_count() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return g_sms;
}

// one launch packs every 128x128 block a GMP call needs; returns the byte stride between blocks
size_t gmp_pack_stride(int mode) { return (size_t)(mode == BSMS_MODE_FP16X3 ? 2 : 1) * kWBlk; }

int gmp_pack_blocks(const PackList& pl, int mode, uint8_t* out, cudaStream_t st) {
  ProfScope ps_(PK_OTHER, st);
  if (mode == BSMS_MODE_FP16X3)
    k_pack_weights<2><<<pl.n, 256, 0, st>>>(pl, out);
  else
    k_pack_weights<1><<<pl.n, 256, 0, st>>>(pl, out);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

int lin_tc(int mode, const float* X0, int ldx0, const float* X1, int ldx1, int KB, int NB, const uint8_t* const* blocks,
           int b_mn, const float* bias, int relu, const float* mask, int ldmask, int accum, float* Y, int ldy,
           long long rows, int kind, cudaStream_t st) {
  if (rows == 0) return BSMS_OK;
  LinParams p;
  p.X[0] = X0;
  p.X[1] = X1;
  p.ldx[0] = ldx0;
  p.ldx[1] = ldx1;
  p.KB = KB;
  p.NB = NB;
  for (int i = 0; i < 4; ++i) p.wblk[i] = nullptr;
  for (int nb = 0; nb < NB; ++nb)
    for (int kb = 0; kb < KB; ++kb) p.wblk[nb * 2 + kb] = blocks[nb * KB + kb];
  p.b_mn = b_mn;
  p.bias = bias;
  p.relu = relu;
  p.mask = mask;
  p.ldmask = ldmask;
  p.accum = accum;
  p.Y = Y;
  p.ldy = ldy;
  p.rows = rows;
  p.ntiles = ceil_div(rows, 128);
  const int grid = std::min(2 * sm_count(), p.ntiles);
  ProfScope ps_(kind, st);
  if (mode == BSMS_MODE_FP16X3) {
    const size_t smem = 1024 + 4 * kWBlk + 1024 + 64;  // 2 blocks of 64 KB: one CTA per SM by shared memory
    BSMS_CUDA(cudaFuncSetAttribute(k_lin<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_lin<2><<<grid, 128, smem, st>>>(p);
  } else {
    // 2 blocks of 32 KB; padded to 80 KB so that at most two CTAs (2 x 256 TMEM columns) share an SM
    const size_t smem = 80 * 1024;
    BSMS_CUDA(cudaFuncSetAttribute(k_lin<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_lin<1><<<grid, 128, smem, st>>>(p);
  }
  BSMS_LAUNCHED();
  return BSMS_OK;
}


// bf16 layer with row-cooperative I/O (k_lin2).  add0/add1: optional addends of the two output blocks; Y0/Y1
// separate destinations; ln_out: fused LayerNorm + residual (NB == 1).
int lin_tc2(const float* X0, int ldx0, const float* X1, int ldx1, int KB, int NB, const uint8_t* const* blocks, int b_mn,
            const float* bias, int relu, const float* mask, int ldmask, const float* add0, int ldadd0, const float* add1,
            int ldadd1, float* Y0, int ldy0, float* Y1, int ldy1, float* ln_out, const float* res0, const float* res1,
            long long rows, int kind, cudaStream_t st) {
  if (rows == 0) return BSMS_OK;
  Lin2Params p;
  p.X[0] = X0; p.X[1] = X1;
  p.ldx[0] = ldx0; p.ldx[1] = ldx1;
  p.KB = KB; p.NB = NB;
  for (int i = 0; i < 4; ++i) p.wblk[i] = nullptr;
  for (int nb = 0; nb < NB; ++nb)
    for (int kb = 0; kb < KB; ++kb) p.wblk[nb * 2 + kb] = blocks[nb * KB + kb];
  p.b_mn = b_mn;
  p.bias = bias;
  p.relu = relu;
  p.mask = mask;
  p.ldmask = ldmask;
  p.add[0] = add0; p.add[1] = add1;
  p.ldadd[0] = ldadd0; p.ldadd[1] = ldadd1;
  p.Y[0] = Y0; p.Y[1] = Y1;
  p.ldy[0] = ldy0; p.ldy[1] = ldy1;
  p.ln_out = ln_out;
  p.res0 = res0;
  p.res1 = res1;
  p.a_scale = 1.f;
  p.a_amax_dev = nullptr;
  p.y_amax_dev = nullptr;
  p.rows = rows;
  p.ntiles = ceil_div(rows, 128);
  const int nblk = KB * NB;
  const int per_sm = nblk == 1 ? 2 : 1;
  const int grid = std::min(per_sm * sm_count(), p.ntiles);
  const size_t smem = 1024 + (2 + nblk) * kWBlk + 1024 + 64;
  BSMS_CUDA(cudaFuncSetAttribute(k_lin2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + 4 * kWBlk + 1024 + 64)));
  ProfScope ps_(kind, st);
  k_lin2<0><<<grid, 256, smem, st>>>(p);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

// The same layer in the two-way fp16 split arithmetic (22 significant bits per operand, three MMAs per K step): the
// tensor-core form of the fp32-parity mode's backward GEMMs.  blocks: the [hi | lo] pairs the forward packed
// (k_pack_weights<2>); a_amax_dev: device slot holding the float bits of max |A| (a gradient operand: its scale is derived
// from it), null = activations (kActScale); y_amax_dev: slot that receives max |output| (block 0) for the next GEMM;
// K = 128 only (NB = 1 or 2).
int lin_tc2_split(const float* X0, int ldx0, int NB, const uint8_t* const* blocks, int b_mn, const float* bias, int relu,
                  const float* mask, int ldmask, const float* add0, int ldadd0, const float* add1, int ldadd1, float* Y0,
                  int ldy0, float* Y1, int ldy1, long long rows, int kind, const unsigned* a_amax_dev, unsigned* y_amax_dev,
                  cudaStream_t st) {
  if (rows == 0) return BSMS_OK;
  Lin2Params p;
  p.X[0] = X0; p.X[1] = nullptr;
  p.ldx[0] = ldx0; p.ldx[1] = 0;
  p.KB = 1; p.NB = NB;
  for (int i = 0; i < 4; ++i) p.wblk[i] = nullptr;
  for (int nb = 0; nb < NB; ++nb) p.wblk[nb * 2] = blocks[nb];
  p.b_mn = b_mn;
  p.bias = bias;
  p.relu = relu;
  p.mask = mask;
  p.ldmask = ldmask;
  p.add[0] = add0; p.add[1] = add1;
  p.ldadd[0] = ldadd0; p.ldadd[1] = ldadd1;
  p.Y[0] = Y0; p.Y[1] = Y1;
  p.ldy[0] = ldy0; p.ldy[1] = ldy1;
  p.ln_out = nullptr;
  p.res0 = p.res1 = nullptr;
  p.a_scale = kActScale;
  p.a_amax_dev = a_amax_dev;
  p.y_amax_dev = y_amax_dev;
  p.rows = rows;
  p.ntiles = ceil_div(rows, 128);
  const int grid = std::min(sm_count(), p.ntiles);
  const size_t smem = 1024 + (2 + 2 * NB) * kWBlk + 1024 + 64;  // A hi | A lo (= the fp32 staging), NB x [W hi | W lo]
  BSMS_CUDA(cudaFuncSetAttribute(k_lin2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(1024 + 6 * kWBlk + 1024 + 64)));
  ProfScope ps_(kind, st);
  k_lin2<1><<<grid, 256, smem, st>>>(p);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

// ------------------------------------------------------------------------------------------
// dW[128,128] += G^T X over all rows with BOTH operands as two-way fp16 splits of scaled values (three MMAs per K
// step, 22 significant bits): the weight gradient of the fp32-parity mode on tensor cores.  One 128 KB stage (G hi,
// G lo, X hi, X lo); G is scaled by the device-resident gradient scale, X by kActScale, the accumulator is scaled back
// when it is flushed; the bias gradient is the exact fp32 column sum of the rows as they pass through the registers.
__global__ void __launch_bounds__(256, 1) k_wgrad_tc_split(const WgradBatch batch, const unsigned* __restrict__ g_amax_dev) {
  extern __shared__ uint8_t smem_raw[];
  const int cpp = batch.ctas_per_prob;
  const WgradParams p = batch.prob[blockIdx.x / cpp];
  const int cta_in_prob = blockIdx.x % cpp;
  constexpr uint32_t IDESC_MM = make_idesc(0, 128, 128, 1, 1);  // fp16 operands
  const float g_scale = split_scale(g_amax_dev[blockIdx.x / batch.ctas_per_prob]);  // one slot per problem
  const float out_scale = 1.f / (g_scale * kActScale);
  const uint32_t s0 = smem_u32(smem_raw);
  const uint32_t sbase = (s0 + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sbase - s0);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(sp + 4 * kWBlk);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);
  const int tid = threadIdx.x, warp = (int)uniform(threadIdx.x >> 5), lane = tid & 31;
  const uint32_t bar0 = smem_u32(&s_bar[0]);
  if (tid == 0) {
    mbar_init(bar0, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 128);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = uniform(*s_tmem);
  uint8_t *t_gh = sp, *t_gl = sp + kWBlk, *t_xh = sp + 2 * kWBlk, *t_xl = sp + 3 * kWBlk;
  float4 acc_b = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t ph = 0;
  int it = 0;
  for (int tile = cta_in_prob; tile < p.ntiles; tile += cpp, ++it) {
    const long long row0 = (long long)tile * 128;
    // the rows of this tile travel to registers while the previous tile's MMAs still read the stage
    float4 vg[8], vx[8];
    coop_rows_load<8>(p.G, p.ldg, row0, p.rows, warp * 16, lane, vg);
    coop_rows_load<8>(p.X, p.ldx, row0, p.rows, warp * 16, lane, vx);
    if (it > 0) {
      mbar_wait(bar0, ph);
      ph ^= 1;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (half == 1) {
        coop_rows_load<8>(p.G, p.ldg, row0, p.rows, warp * 16 + 8, lane, vg);
        coop_rows_load<8>(p.X, p.ldx, row0, p.rows, warp * 16 + 8, lane, vx);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc_b.x += vg[u].x; acc_b.y += vg[u].y; acc_b.z += vg[u].z; acc_b.w += vg[u].w;
      }
      coop_rows_store_split<8>(t_gh, t_gl, warp * 16 + 8 * half, lane, vg, g_scale);
      coop_rows_store_split<8>(t_xh, t_xl, warp * 16 + 8 * half, lane, vx, kActScale);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0) {
      fence_after_sync();
      if (elect_one()) {
        const uint32_t gh = sbase, gl = sbase + kWBlk, xh = sbase + 2 * kWBlk, xl = sbase + 3 * kWBlk;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t dgh = smem_desc_sw128(gh + ks * 2048, 16384, 1024), dgl = smem_desc_sw128(gl + ks * 2048, 16384, 1024);
          const uint64_t dxh = smem_desc_sw128(xh + ks * 2048, 16384, 1024), dxl = smem_desc_sw128(xl + ks * 2048, 16384, 1024);
          mma_ss(tmem_base, dgh, dxh, IDESC_MM, (it > 0 || ks > 0) ? 1u : 0u);
          mma_ss(tmem_base, dgl, dxh, IDESC_MM, 1u);
          mma_ss(tmem_base, dgh, dxl, IDESC_MM, 1u);
        }
        mma_commit(bar0);
      }
      __syncwarp();
    }
  }
  if (it > 0) {
    mbar_wait(bar0, ph);
    fence_after_sync();
    const int q = warp & 3, hh = warp >> 2;  // TMEM lane = output channel n
    const int n = q * 32 + lane;
    uint32_t r0[32], r1[32];
    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + 64 * hh;
    tmem_ld32(ta, r0);
    tmem_ld32(ta + 32, r1);
    wait_ld();
    float* dst = p.dW + (size_t)n * p.ldo + 64 * hh;
#pragma unroll
    for (int t = 0; t < 32; ++t) atomicAdd(dst + t, __uint_as_float(r0[t]) * out_scale);
#pragma unroll
    for (int t = 0; t < 32; ++t) atomicAdd(dst + 32 + t, __uint_as_float(r1[t]) * out_scale);
    if (p.db) {
      atomicAdd(p.db + 4 * lane + 0, acc_b.x);
      atomicAdd(p.db + 4 * lane + 1, acc_b.y);
      atomicAdd(p.db + 4 * lane + 2, acc_b.z);
      atomicAdd(p.db + 4 * lane + 3, acc_b.w);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

WgradParams wgrad_problem(const float* G, int ldg, const float* X, int ldx, float* dW, int ldo, float* db, long long rows);
// plain-pointer front ends for gmp.cu (which does not see the packed-operand headers)
// g_amax_dev: nprob CONSECUTIVE device slots, slot i = float bits of max |G_i|
int wgrad_tc_split_batch(const WgradParams* probs, int nprob, const unsigned* g_amax_dev, cudaStream_t st);
int wgrad_tc_split3(const float* const* G, const int* ldg, const float* const* X, const int* ldx, float* const* dW, const int* ldo,
                    float* const* db, int nprob, long long rows, const unsigned* g_amax_dev, cudaStream_t st) {
  WgradParams pr[6];
  for (int i = 0; i < nprob; ++i) pr[i] = wgrad_problem(G[i], ldg[i], X[i], ldx[i], dW[i], ldo[i], db[i], rows);
  return wgrad_tc_split_batch(pr, nprob, g_amax_dev, st);
}

int wgrad_tc_split_batch(const WgradParams* probs, int nprob, const unsigned* g_amax_dev, cudaStream_t st) {
  if (nprob == 0 || probs[0].rows == 0) return BSMS_OK;
  WgradBatch b;
  b.nprob = nprob;
  const int ntiles = ceil_div(probs[0].rows, 128);
  for (int i = 0; i < nprob; ++i) {
    b.prob[i] = probs[i];
    b.prob[i].ntiles = ntiles;
  }
  b.ctas_per_prob = std::max(1, std::min(ntiles, sm_count() / nprob));
  const size_t smem = 1024 + 4 * kWBlk + 64;
  BSMS_CUDA(cudaFuncSetAttribute(k_wgrad_tc_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps_(PK_WGRAD, st);
  k_wgrad_tc_split<<<b.ctas_per_prob * nprob, 256, smem, st>>>(b, g_amax_dev);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

// One launch for up to 6 weight-gradient problems over the same number of rows.
int wgrad_tc_batch(const WgradParams* probs, int nprob, cudaStream_t st, float* part) {
  if (nprob == 0 || probs[0].rows == 0) return BSMS_OK;
  WgradBatch b;
  b.nprob = nprob;
  b.part = part;
  const int ntiles = ceil_div(probs[0].rows, 128);
  for (int i = 0; i < nprob; ++i) {
    b.prob[i] = probs[i];
    b.prob[i].ntiles = ntiles;
  }
  b.ctas_per_prob = std::max(1, std::min(ntiles, sm_count() / nprob));
  const size_t smem = 1024 + 4 * kWBlk + 64;
  BSMS_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps_(PK_WGRAD, st);
  k_wgrad_tc<<<b.ctas_per_prob * nprob, 256, smem, st>>>(b);
  BSMS_LAUNCHED();
  if (part) {
    DetSeg segs[12];
    int ns = 0;
    for (int i = 0; i < nprob; ++i) {
      segs[ns++] = DetSeg{probs[i].dW, 0, i * b.ctas_per_prob, b.ctas_per_prob, 1, 128, 128, 128, probs[i].ldo};
      if (probs[i].db) segs[ns++] = DetSeg{probs[i].db, 16384, i * b.ctas_per_prob, b.ctas_per_prob, 2, 1, 128, 128, 128};
    }
    return det_reduce(part, kDetWgradStride, segs, ns, st);
  }
  return BSMS_OK;
}

WgradParams wgrad_problem(const float* G, int ldg, const float* X, int ldx, float* dW, int ldo, float* db, long long rows) {
  WgradParams p;
  p.G = G;
  p.ldg = ldg;
  p.X = X;
  p.ldx = ldx;
  p.dW = dW;
  p.ldo = ldo;
  p.db = db;
  p.rows = rows;
  p.ntiles = 0;
  return p;
}

int wgrad_tc(const float* G, int ldg, const float* X, int ldx, float* dW, int ldo, float* db, long long rows,
             cudaStream_t st) {
  WgradParams p = wgrad_problem(G, ldg, X, ldx, dW, ldo, db, rows);
  return wgrad_tc_batch(&p, 1, st, nullptr);
}

// Test hook: Y = ((X W^T) or (X W)) [. (mask > 0)] on the split-operand tensor-core layer; a_is_grad: scale X from its
// own max |value| (computed here), else the static activation scale.  scratch: 64 KB + 64 bytes.
__global__ void k_dbg_amax(const float* __restrict__ g, long long n, unsigned* __restrict__ slot) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(g[i]));
  bsms::publish_amax(slot, m);
}
}  // namespace bsms
extern "C" int bsms_debug_lin_split(const float* X, int64_t rows, const float* W, int32_t b_mn, const float* mask,
                                    int32_t a_is_grad, float* Y, void* scratch, void* stream) {
  using namespace bsms;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* pack = (uint8_t*)scratch;
  unsigned* slot = (unsigned*)(pack + 2 * kWBlk);
  PackList pl;
  pl.n = 1;
  pl.w[0] = W;
  pl.ld[0] = kD;
  k_pack_weights<2><<<1, 256, 0, st>>>(pl, pack);
  BSMS_CUDA(cudaMemsetAsync(slot, 0, 16, st));
  if (a_is_grad) k_dbg_amax<<<64, 256, 0, st>>>(X, rows * kD, slot);
  const uint8_t* b[1] = {pack};
  return lin_tc2_split(X, kD, 1, b, b_mn, nullptr, 0, mask, kD, nullptr, 0, nullptr, 0, Y, kD, nullptr, 0, rows, PK_OTHER,
                       a_is_grad ? slot : nullptr, slot + 1, st);
}
