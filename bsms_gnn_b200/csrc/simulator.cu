// The caller side of the processor for inference / rollouts (reference: src/models/model.py:83-106,127-164,
// src/utils/normalizer.py:39-52,80-83, src/utils/rollout_utils.py:48-62), fused around it:
//   k_encode_in   : split node_in into [state | pos | type], normalise [state, type] with the input normaliser
//                   (fp64 like the reference, cast to fp32), first encoder Linear (K = out_dim + 1 <= 8) + ReLU,
//                   and the contiguous position tensor the processor consumes — one launch, one pass;
//   dense128_stack: the 128 -> 128 layers of the encoder / decoder MLPs (src/ops/basic.py:6-23) on the node-level
//                   GEMM kernels of the selected arithmetic mode (tcgen05 in the tensor-core modes);
//   k_decode_out  : last decoder Linear (N = out_dim <= 4), inverse target normalisation, mask, residual to the
//                   state, and — for rollouts — the next input row with the boundary nodes re-imposed from the
//                   initial condition (rollout_utils.py:57-62), again one launch.
#include "chain.cuh"

namespace bsms {

constexpr int kMaxIn = 8;   // encoder input width out_dim + 1
constexpr int kMaxOut = 4;  // decoder output width out_dim

struct EncodeParams {
  const float* node_in;  // [rows, Cin], Cin = C + P + 1: [state(C) | pos(P) | type(1)]
  int Cin, C, P;
  double mean[kMaxIn], inv_std[kMaxIn];  // input normaliser over [state, type]
  const float* W0;  // [128, C + 1]
  const float* b0;  // [128]
  float* a1;        // [rows, 128] relu(W0 xn + b0)
  float* pos;       // [rows, P]
  long long rows;
};

// one warp per row, lane l owns output channels 4l..4l+3
__global__ void __launch_bounds__(256) k_encode_in(const EncodeParams p) {
  const int lane = threadIdx.x & 31;
  const int K = p.C + 1;
  float w[4][kMaxIn], b[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    b[c] = p.b0[4 * lane + c];
#pragma unroll
    for (int k = 0; k < kMaxIn; ++k) w[c][k] = k < K ? p.W0[(size_t)(4 * lane + c) * K + k] : 0.f;
  }
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < p.rows; r += ((long long)gridDim.x * blockDim.x) >> 5) {
    const float* row = p.node_in + r * p.Cin;
    float xn[kMaxIn];
#pragma unroll
    for (int k = 0; k < kMaxIn; ++k) {
      // [state | type]: channel k < C is state k, channel C is the node type (model.py:29-46)
      const float raw = k < p.C ? row[k] : (k == p.C ? row[p.Cin - 1] : 0.f);
      xn[k] = k < K ? (float)(((double)raw - p.mean[k]) * p.inv_std[k]) : 0.f;
    }
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxIn; ++k) acc = fmaf(xn[k], w[c][k], acc);
      o[c] = fmaxf(acc + b[c], 0.f);
    }
    st4(p.a1 + r * 128 + 4 * lane, make_float4(o[0], o[1], o[2], o[3]));
    if (lane < p.P) p.pos[r * p.P + lane] = row[p.C + lane];
  }
}

struct DecodeParams {
  const float* y;        // [rows, 128] decoder activation after its three ReLU layers
  const float* W3;       // [C, 128]
  const float* b3;       // [C]
  double mean[kMaxOut], std[kMaxOut];  // target normaliser
  const float* node_in;  // [rows, Cin] the input this step started from (state = its first C channels)
  const float* mask;     // [rows]
  const float* ic;       // [rows, Cin] initial condition for the boundary re-imposition, or null
  float* pred;           // [rows, C]
  float* next_in;        // [rows, Cin] or null
  int Cin, C;
  int pos_feedback;  // next_in's position channels receive the new state (deforming meshes, C == P)
  long long rows;
};

__global__ void __launch_bounds__(256) k_decode_out(const DecodeParams p) {
  const int lane = threadIdx.x & 31;
  float w[kMaxOut][4];
#pragma unroll
  for (int c = 0; c < kMaxOut; ++c) {
    const float4 v = c < p.C ? ld4(p.W3 + (size_t)c * 128 + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    w[c][0] = v.x; w[c][1] = v.y; w[c][2] = v.z; w[c][3] = v.w;
  }
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < p.rows; r += ((long long)gridDim.x * blockDim.x) >> 5) {
    const float4 yv = ld4(p.y + r * 128 + 4 * lane);
    float o[kMaxOut];
#pragma unroll
    for (int c = 0; c < kMaxOut; ++c) o[c] = warp_sum(yv.x * w[c][0] + yv.y * w[c][1] + yv.z * w[c][2] + yv.w * w[c][3]);
    const float* in = p.node_in + r * p.Cin;
    const float mk = p.mask[r];
    if (lane < p.Cin) {
      float v = in[lane];  // pos / type channels pass through
      const int ci = lane < p.C ? lane : ((p.pos_feedback && lane < 2 * p.C) ? lane - p.C : -1);
      if (ci >= 0) {
        float oc = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxOut; ++c) oc = ci == c ? o[c] : oc;
        // Normalizer.inverse (normalizer.py:80-83): fp64 product, cast to fp32; then mask and residual (model.py:158-163)
        const float delta = (float)((double)(oc + p.b3[ci]) * p.std[ci] + p.mean[ci]);
        v = in[ci] + delta * mk;
        if (lane < p.C) p.pred[r * p.C + lane] = v;
      }
      if (p.next_in) p.next_in[r * p.Cin + lane] = (p.ic && mk == 0.f) ? p.ic[r * p.Cin + lane] : v;  // rollout_utils.py:57-62
    }
  }
}

// kernels / launchers of the node-level GEMMs (node_gemm.cu, gmp.cu)
size_t gmp_pack_stride(int mode);
int gmp_pack_blocks(const PackList& pl, int mode, uint8_t* out, cudaStream_t st);
int lin_tc(int mode, const float* X0, int ldx0, const float* X1, int ldx1, int KB, int NB, const uint8_t* const* blocks,
           int b_mn, const float* bias, int relu, const float* mask, int ldmask, int accum, float* Y, int ldy,
           long long rows, int kind, cudaStream_t st);
int lin_tc2(const float* X0, int ldx0, const float* X1, int ldx1, int KB, int NB, const uint8_t* const* blocks, int b_mn,
            const float* bias, int relu, const float* mask, int ldmask, const float* add0, int ldadd0, const float* add1,
            int ldadd1, float* Y0, int ldy0, float* Y1, int ldy1, float* ln_out, const float* res0, const float* res1,
            long long rows, int kind, cudaStream_t st);
int launch_ln_residual(const float* Yn, const float* x, const float* skip, float* out, long long rows, cudaStream_t st);
int fp32_linear128(const float* X, const float* W, const float* bias, int relu, float* Y, long long rows, cudaStream_t st);

}  // namespace bsms

using namespace bsms;

#define SIM_TRY(expr)             \
  do {                            \
    int _rc = (expr);             \
    if (_rc != BSMS_OK) return _rc; \
  } while (0)

extern "C" int bsms_encode_in(const float* node_in, int64_t rows, int32_t Cin, int32_t C, int32_t P, const double* mean_host,
                              const double* std_host, const float* W0, const float* b0, float* a1, float* pos, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CHECK_ARG(node_in && mean_host && std_host && W0 && b0 && a1 && pos && rows >= 1, "bsms_encode_in: null argument");
  BSMS_CHECK_ARG(C >= 1 && C + 1 <= kMaxIn && P >= 1 && P <= 3 && Cin == C + P + 1, "bsms_encode_in: need Cin = C + P + 1, C + 1 <= %d", kMaxIn);
  EncodeParams p;
  p.node_in = node_in;
  p.Cin = Cin;
  p.C = C;
  p.P = P;
  for (int k = 0; k < kMaxIn; ++k) {
    p.mean[k] = k <= C ? mean_host[k] : 0.0;
    p.inv_std[k] = k <= C ? 1.0 / std_host[k] : 0.0;
  }
  p.W0 = W0;
  p.b0 = b0;
  p.a1 = a1;
  p.pos = pos;
  p.rows = rows;
  ProfScope ps_(PK_OTHER, st);
  k_encode_in<<<(int)std::min<long long>(ceil_div(rows * 32, 256), 148 * 8), 256, 0, st>>>(p);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

extern "C" int bsms_decode_out(const float* y, int64_t rows, int32_t Cin, int32_t C, const float* W3, const float* b3,
                               const double* mean_host, const double* std_host, const float* node_in, const float* mask,
                               const float* ic, float* pred, float* next_in, int32_t pos_feedback, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CHECK_ARG(y && W3 && b3 && mean_host && std_host && node_in && mask && pred && rows >= 1, "bsms_decode_out: null argument");
  BSMS_CHECK_ARG(C >= 1 && C <= kMaxOut && Cin > C && Cin <= 32, "bsms_decode_out: out_dim %d unsupported (1..%d)", C, kMaxOut);
  DecodeParams p;
  p.y = y;
  p.W3 = W3;
  p.b3 = b3;
  for (int c = 0; c < kMaxOut; ++c) {
    p.mean[c] = c < C ? mean_host[c] : 0.0;
    p.std[c] = c < C ? std_host[c] : 0.0;
  }
  p.node_in = node_in;
  p.mask = mask;
  p.ic = ic;
  p.pred = pred;
  p.next_in = next_in;
  p.pos_feedback = pos_feedback;
  BSMS_CHECK_ARG(!pos_feedback || Cin == 2 * C + 1, "bsms_decode_out: pos_feedback needs pos_dim == out_dim");
  p.Cin = Cin;
  p.C = C;
  p.rows = rows;
  ProfScope ps_(PK_OTHER, st);
  k_decode_out<<<(int)std::min<long long>(ceil_div(rows * 32, 256), 148 * 8), 256, 0, st>>>(p);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

// packed images of up to three 128x128 layers: n_layers x stride bytes (bsms_dense128_packed_bytes)
extern "C" size_t bsms_dense128_packed_bytes(int32_t mode) { return mode == BSMS_MODE_FP32 ? 256 : 3 * gmp_pack_stride(mode); }

extern "C" int bsms_dense128_pack(const float* const* W_host3, int32_t n_layers, int32_t mode, void* packed, void* stream) {
  BSMS_CHECK_ARG(W_host3 && packed && n_layers >= 1 && n_layers <= 3, "bsms_dense128_pack: 1..3 layers");
  if (mode == BSMS_MODE_FP32) return BSMS_OK;
  PackList pl;
  pl.n = n_layers;
  for (int l = 0; l < n_layers; ++l) {
    pl.w[l] = W_host3[l];
    pl.ld[l] = kD;
  }
  return gmp_pack_blocks(pl, mode, (uint8_t*)packed, (cudaStream_t)stream);
}

// x [rows,128] -> n_layers x (Linear 128->128 [+ ReLU where bit l of relu_mask is set]) [-> LayerNorm] -> out [rows,128].
// scratch: 2 x rows x 128 floats.  `packed` from bsms_dense128_pack (ignored in fp32 mode).
extern "C" int bsms_dense128_stack(const float* x, int64_t rows, const float* const* W_host3, const float* const* b_host3,
                                   int32_t n_layers, int32_t relu_mask, int32_t layer_norm, int32_t mode, const void* packed,
                                   float* out, float* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CHECK_ARG(x && W_host3 && b_host3 && out && scratch && rows >= 1 && n_layers >= 1 && n_layers <= 3,
                 "bsms_dense128_stack: bad argument");
  BSMS_CHECK_ARG(mode == BSMS_MODE_FP32 || packed, "bsms_dense128_stack: the tensor-core modes need packed weights");
  float* buf[2] = {scratch, scratch + rows * kD};
  const float* cur = x;
  const size_t bs = mode == BSMS_MODE_FP32 ? 0 : gmp_pack_stride(mode);
  for (int l = 0; l < n_layers; ++l) {
    const bool last = l == n_layers - 1;
    const int relu = (relu_mask >> l) & 1;
    float* dst = (last && !layer_norm) ? out : buf[l & 1];
    if (mode == BSMS_MODE_FP32) {
      SIM_TRY(fp32_linear128(cur, W_host3[l], b_host3[l], relu, dst, rows, st));
    } else {
      const uint8_t* blk[1] = {(const uint8_t*)packed + (size_t)l * bs};
      if (mode == BSMS_MODE_BF16) {
        // the last layer fuses the LayerNorm into its store pass (no residual)
        SIM_TRY(lin_tc2(cur, kD, nullptr, 0, 1, 1, blk, 0, b_host3[l], relu, nullptr, 0, nullptr, 0, nullptr, 0, dst, kD, nullptr, 0,
                        (last && layer_norm) ? out : nullptr, nullptr, nullptr, rows, PK_NODE_FWD_GEMM, st));
      } else {
        SIM_TRY(lin_tc(mode, cur, kD, nullptr, 0, 1, 1, blk, 0, b_host3[l], relu, nullptr, 0, 0, dst, kD, rows, PK_NODE_FWD_GEMM, st));
      }
    }
    cur = dst;
  }
  if (layer_norm && mode != BSMS_MODE_BF16) SIM_TRY(launch_ln_residual(cur, nullptr, nullptr, out, rows, st));
  return BSMS_OK;
}
