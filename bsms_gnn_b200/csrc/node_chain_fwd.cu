// Fused forward of the node MLP's layers 1..3 on tcgen05 (bf16 operands, fp32 accumulate):
//   N2 = relu(N1 V2^T + c2) ; N3 = relu(N2 V3^T + c3) ; Yn = N3 V4^T + c4 ; out = LN(Yn) + x (+ skip)
// in ONE persistent kernel (reference: src/ops/basic.py:6-23,97-98 and BSMS.py:102).  N1 comes in as
// fp32 rows (row-cooperative loads, the next tile's rows are prefetched into registers under the MMAs);
// the activations N2 / N3 never leave the chip as fp32: each epilogue writes them as a bf16 UMMA operand
// tile in shared memory, where the next layer's SS-form MMA reads them, and the SAME tile image goes to
// HBM with one cp.async.bulk store (32 KB per tile) for the backward kernel, which loads it back with
// cp.async.bulk straight into operand position (node_chain_bwd.cu) — no conversion on either side, half
// the bytes of an fp32 row.  Only Yn (needed in fp32 by the LayerNorm backward) and the block output
// pass through the fp32 staging tile and leave as coalesced 512 B rows.
#include "chain.cuh"

namespace bsms {

struct NodeFwdParams {
  const float* N1;       // [rows,128] activation of the first node layer
  const uint8_t* wpack;  // packed bf16 blocks V2, V3, V4 (contiguous)
  const float* bias[3];  // c2, c3, c4
  uint8_t* img[2];       // N2, N3 as bf16 operand-tile images: ntiles x 32 KB each
  float* Yn;             // [rows,128] pre-LayerNorm output (kept for backward)
  const float* x;        // residual input of the block
  const float* skip;     // optional second residual (U-Net skip), may be null
  float* out;            // [rows,128] LN(Yn) + x (+ skip); null: only the kept tensors are produced
  long long rows;
  int ntiles;
};

__device__ __forceinline__ uint32_t nf_off(int r, int chunk) {  // 16-byte chunk `chunk` (0..15) of tile row r
  return (uint32_t)((chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(256, 1) k_node_chain_fwd(const NodeFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t IDESC = make_idesc(1, 128, 128, 0, 0);  // A, B K-major
  const uint32_t s0 = smem_u32(smem_raw);
  const uint32_t sbase = (s0 + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sbase - s0);
  // slots: [V2][V3][V4][TA][TB]; the fp32 staging of Yn aliases TA|TB
  uint8_t* s_TA = sp + 3 * kWBlk;
  uint8_t* s_TB = sp + 4 * kWBlk;
  float* s_stage = reinterpret_cast<float*>(s_TA);
  float* s_bias = reinterpret_cast<float*>(sp + 5 * kWBlk);  // [3][128]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 384);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2);
  const uint32_t aV[3] = {sbase, sbase + kWBlk, sbase + 2 * kWBlk};
  const uint32_t aTA = sbase + 3 * kWBlk, aTB = sbase + 4 * kWBlk;

  const int tid = threadIdx.x, warp = (int)uniform(threadIdx.x >> 5), lane = tid & 31;
  const int q = warp & 3, h = warp >> 2, r = q * 32 + lane;
  const uint32_t bar_w = smem_u32(&s_bar[0]), bar_m = smem_u32(&s_bar[1]);
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_m, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 128);
  for (int i = tid; i < 384; i += 256) s_bias[i] = p.bias[i >> 7][i & 127];
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t d_tmem = uniform(*s_tmem);
  if (tid == 0) {
    mbar_expect_tx(bar_w, 3 * kWBlk);
    for (int blk = 0; blk < 3; ++blk) bulk_g2s(aV[blk], p.wpack + (size_t)blk * kWBlk, kWBlk, bar_w);
    mbar_wait(bar_w, 0);
  }
  __syncwarp();  // lane 0 rejoins its warp (a warp left split runs its collectives on the slow path until the next barrier)
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  const uint32_t d_mine = d_tmem + lane_off + 64 * h;
  uint32_t phase = 0;

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
  // D = A(tile, K-major) x W(block, K-major)^T
  auto issue = [&](uint32_t a_tile, uint32_t w_blk) {
    if (warp == 0) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t koff = (ks >> 2) * 16384 + (ks & 3) * 32;
          mma_ss(d_tmem, smem_desc_sw128(a_tile + koff, 16, 1024), smem_desc_sw128(w_blk + koff, 16, 1024), IDESC, ks > 0);
        }
        mma_commit(bar_m);
      }
      __syncwarp();
    }
  };
  // activation epilogue: relu(D + bias) -> bf16 operand tile (thread = row r, column half h)
  auto act_epilogue = [&](const float* bias, uint8_t* dst_tile) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t rr_[32];
      tmem_ld32(d_mine + 32 * hh, rr_);
      wait_ld();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = fmaxf(__uint_as_float(rr_[8 * jj + e]) + bias[64 * h + 32 * hh + 8 * jj + e], 0.f);
        uint4 u;
        u.x = pack_bf16(o8[0], o8[1]); u.y = pack_bf16(o8[2], o8[3]);
        u.z = pack_bf16(o8[4], o8[5]); u.w = pack_bf16(o8[6], o8[7]);
        *reinterpret_cast<uint4*>(dst_tile + nf_off(r, 8 * h + 4 * hh + jj)) = u;
      }
    }
  };

  float4 pre[16];  // the next tile's N1 rows 16*warp .. 16*warp+15
  if ((int)blockIdx.x < p.ntiles) coop_rows_load<16>(p.N1, kD, (long long)blockIdx.x * 128, p.rows, warp * 16, lane, pre);

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * 128;
    coop_rows_store<16>(s_TA, warp * 16, lane, pre);
    sync_all();
    issue(aTA, aV[0]);  // D = N1 V2^T
    {
      const int next = tile + gridDim.x;
      if (next < p.ntiles) coop_rows_load<16>(p.N1, kD, (long long)next * 128, p.rows, warp * 16, lane, pre);
    }
    wait_mma();
    act_epilogue(s_bias, s_TB);  // N2 -> TB
    sync_all();
    if (tid == 0) bulk_s2g(p.img[0] + (size_t)tile * kWBlk, aTB, kWBlk);  // N2 image -> HBM
    issue(aTB, aV[1]);  // D = N2 V3^T
    wait_mma();
    act_epilogue(s_bias + 128, s_TA);  // N3 -> TA (N1's MMA completed)
    sync_all();
    if (tid == 0) bulk_s2g(p.img[1] + (size_t)tile * kWBlk, aTA, kWBlk);  // N3 image -> HBM
    issue(aTA, aV[2]);  // D = N3 V4^T
    wait_mma();
    if (tid == 0) bulk_wait_read();  // both image stores have read their tiles: TA|TB may be overwritten
    __syncthreads();
    // ---- Yn = D + c4 -> fp32 staging over TA|TB (16-byte chunks XOR-swizzled by row)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t rr_[32];
      tmem_ld32(d_mine + 32 * hh, rr_);
      wait_ld();
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const int c = 64 * h + 32 * hh + 4 * q4;
        const float4 o = make_float4(__uint_as_float(rr_[4 * q4 + 0]) + s_bias[256 + c + 0], __uint_as_float(rr_[4 * q4 + 1]) + s_bias[256 + c + 1],
                                     __uint_as_float(rr_[4 * q4 + 2]) + s_bias[256 + c + 2], __uint_as_float(rr_[4 * q4 + 3]) + s_bias[256 + c + 3]);
        *reinterpret_cast<float4*>(s_stage + r * 128 + (((c >> 2) ^ (r & 31)) << 2)) = o;
      }
    }
    __syncthreads();
    // ---- row-cooperative pass: Yn rows out, LayerNorm + residual(s) -> block output.  The residual rows of
    //      8 tile rows are requested together (16 loads in flight per warp) before the rows are normalised:
    //      one exposed DRAM latency per 8 rows instead of one per row
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int rb = warp * 16 + 8 * half;
      float4 xr[8], sk[8];
      if (p.out) {
        coop_rows_load<8>(p.x, kD, row0, p.rows, rb, lane, xr);
        if (p.skip) coop_rows_load<8>(p.skip, kD, row0, p.rows, rb, lane, sk);
      }
      // four rows at a time: their LayerNorm sums (one pass over d = y - shift, shift = the row's first element: no
      // cancellation in E[d^2] - E[d]^2) share one butterfly, 8 independent shuffles per step instead of two dependent
      // 5-step reductions per row
#pragma unroll
      for (int u0 = 0; u0 < 8; u0 += 4) {
        float4 v[4];
        float sh[4], s1[4], s2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rr = rb + u0 + k;
          v[k] = *reinterpret_cast<const float4*>(s_stage + rr * 128 + ((lane ^ (rr & 31)) << 2));
          if (row0 + rr < p.rows) st4(p.Yn + (row0 + rr) * kD + 4 * lane, v[k]);
        }
        if (p.out) {
#pragma unroll
          for (int k = 0; k < 4; ++k) sh[k] = __shfl_sync(0xffffffffu, v[k].x, 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float dx = v[k].x - sh[k], dy = v[k].y - sh[k], dz = v[k].z - sh[k], dw = v[k].w - sh[k];
            s1[k] = (dx + dy) + (dz + dw);
            s2[k] = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            float t1[4], t2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              t1[k] = __shfl_xor_sync(0xffffffffu, s1[k], o);
              t2[k] = __shfl_xor_sync(0xffffffffu, s2[k], o);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              s1[k] += t1[k];
              s2[k] += t2[k];
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rr = rb + u0 + k;
            const long long row = row0 + rr;
            const float md = s1[k] * (1.f / 128.f);
            const float var = fmaxf(s2[k] * (1.f / 128.f) - md * md, 0.f);
            const float rstd = 1.f / sqrtf(var + 1e-5f);
            const float mean = sh[k] + md;
            float4 o = xr[u0 + k];
            o.x += (v[k].x - mean) * rstd; o.y += (v[k].y - mean) * rstd; o.z += (v[k].z - mean) * rstd; o.w += (v[k].w - mean) * rstd;
            if (p.skip) {
              o.x += sk[u0 + k].x; o.y += sk[u0 + k].y; o.z += sk[u0 + k].z; o.w += sk[u0 + k].w;
            }
            if (row < p.rows) st4(p.out + row * kD + 4 * lane, o);
          }
        }
      }
    }
    __syncthreads();  // the staging tile is rewritten (as TA, by generic stores) by the next tile
  }
  if (tid == 0) bulk_wait_all();  // the image stores are complete before the CTA retires
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(d_tmem, 128);
}

// N2 / N3 leave as bf16 operand-tile images (img2 / img3: ceil(rows/128) x 32 KB), Yn and `out` as fp32 rows.
int node_chain_forward(const float* N1, const uint8_t* wpack_v2, const float* c2, const float* c3, const float* c4,
                       uint8_t* img2, uint8_t* img3, float* Yn, const float* x, const float* skip, float* out,
                       long long rows, cudaStream_t st) {
  if (rows == 0) return BSMS_OK;
  NodeFwdParams p;
  p.N1 = N1;
  p.wpack = wpack_v2;
  p.bias[0] = c2;
  p.bias[1] = c3;
  p.bias[2] = c4;
  p.img[0] = img2;
  p.img[1] = img3;
  p.Yn = Yn;
  p.x = x;
  p.skip = skip;
  p.out = out;
  p.rows = rows;
  p.ntiles = ceil_div(rows, 128);
  int dev = 0, sms = 148;
  BSMS_CUDA(cudaGetDevice(&dev));
  BSMS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = 1024 + 5 * kWBlk + 384 * 4 + 2 * 8 + 16;
  BSMS_CUDA(cudaFuncSetAttribute(k_node_chain_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps_(PK_NODE_FWD_GEMM, st);
  k_node_chain_fwd<<<std::min(sms, p.ntiles), 256, smem, st>>>(p);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

}  // namespace bsms
