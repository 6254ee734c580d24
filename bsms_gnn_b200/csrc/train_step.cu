// The rest of the training step around the processor (reference: src/trainer/trainer.py:79-98,134-156,
// src/utils/basic.py:168-184, configs/opt/default.yaml): masked RMSE loss + its gradient, global-norm
// gradient clipping and AdamW with the warmup-cosine learning-rate schedule, as a handful of streaming
// kernels over ONE flat parameter / gradient / moment buffer instead of ~4 ATen launches per parameter
// tensor (the reference's 132 tensors -> ~600 launches per step).  Every scalar that changes from step to
// step (loss sums, gradient norm, step counter, learning rate) lives on the device, so the whole step is
// free of host synchronisation and CUDA-graph capturable.  All kernels are HBM-bound elementwise passes.
#include <curand_kernel.h>
#include <math.h>

#include "common.cuh"

namespace bsms {

// ---- masked RMSE (trainer.py:96-98): rmse = sqrt( sum(se * mask) / sum(mask) / C ), se = (pred - tar)^2,
//      mask [rows] broadcast over the C channels.  acc[0] += sum(se*mask), acc[1] += sum(mask).
__global__ void __launch_bounds__(256)
k_rmse_partial(const float* __restrict__ pred, const float* __restrict__ tar, const float* __restrict__ mask, long long rows,
               int C, double* __restrict__ acc) {
  double s = 0.0, m = 0.0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float mk = mask[r];
    float se = 0.f;
    for (int c = 0; c < C; ++c) {
      const float d = pred[r * C + c] - tar[r * C + c];
      se = fmaf(d, d, se);
    }
    s += (double)(se * mk);
    m += (double)mk;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    m += __shfl_xor_sync(0xffffffffu, m, o);
  }
  __shared__ double sh[2][8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sh[0][w] = s;
    sh[1][w] = m;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) {
      a += sh[0][i];
      b += sh[1][i];
    }
    atomicAdd(acc, a);
    atomicAdd(acc + 1, b);
  }
}
// loss = sqrt(acc0 / acc1 / C); grad_pred = g_loss * mask * (pred - tar) / (loss * acc1 * C)
__global__ void __launch_bounds__(256)
k_rmse_finish(const float* __restrict__ pred, const float* __restrict__ tar, const float* __restrict__ mask, long long rows,
              int C, const double* __restrict__ acc, const float* __restrict__ g_loss, float* __restrict__ loss,
              float* __restrict__ grad_pred) {
  const double ms = acc[1] * (double)C;
  const double l = sqrt(acc[0] / ms);
  if (blockIdx.x == 0 && threadIdx.x == 0 && loss) *loss = (float)l;
  if (!grad_pred) return;
  const float k = (float)((g_loss ? (double)*g_loss : 1.0) / (l * ms));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * C; i += (long long)gridDim.x * blockDim.x)
    grad_pred[i] = k * mask[i / C] * (pred[i] - tar[i]);
}

// ---- sum of squares of the flat gradient (global norm, torch.nn.utils.clip_grad_norm_)
__global__ void __launch_bounds__(256) k_sumsq(const float* __restrict__ g, long long n, double* __restrict__ out) {
  double s = 0.0;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    s += (double)(v * v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += sh[i];
    atomicAdd(out, a);
  }
}

// ---- step scalars on the device.  state[0] = optimiser step t (0-based count of updates done so far, as a double);
//      hyper = {lr_t, 1 - beta1^(t+1), sqrt(1 - beta2^(t+1)), clip coefficient}.
//      lr_t = peak_lr * factor(t), factor = t / warmup for t <= warmup, else 0.5 (1 + cos(pi (t - warmup) / (max - warmup)))
//      (WarmupCosineDecayScheduler, src/utils/basic.py:168-184: the scheduler's epoch equals the number of updates done).
__global__ void k_step_scalars(double* __restrict__ state, const double* __restrict__ gnorm_sq, float* __restrict__ hyper,
                               double peak_lr, double warmup, double max_iters, double beta1, double beta2, double max_norm) {
  const double t = state[0];
  double f;
  if (warmup <= 0.0 && max_iters <= 0.0)
    f = 1.0;  // constant learning rate
  else if (t <= warmup)
    f = warmup > 0.0 ? t / warmup : 1.0;
  else
    f = 0.5 * (1.0 + cos(3.14159265358979323846 * (t - warmup) / (max_iters - warmup)));
  hyper[0] = (float)(peak_lr * f);
  hyper[1] = (float)(1.0 - pow(beta1, t + 1.0));
  hyper[2] = (float)sqrt(1.0 - pow(beta2, t + 1.0));
  double coef = 1.0;
  if (max_norm > 0.0 && gnorm_sq) {
    coef = max_norm / (sqrt(*gnorm_sq) + 1e-6);  // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
    if (coef > 1.0) coef = 1.0;
  }
  hyper[3] = (float)coef;
  state[0] = t + 1.0;
}

// ---- clip + AdamW over the flat buffers (torch.optim.AdamW, amsgrad off):
//      g *= coef; p *= 1 - lr wd; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
k_clip_adamw(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
             const float* __restrict__ hyper, float beta1, float beta2, float eps, float wd, int zero_grad) {
  const float lr = hyper[0], bc1 = hyper[1], sbc2 = hyper[2], coef = hyper[3];
  const float decay = 1.f - lr * wd, step = lr / bc1;
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= coef;
    pp *= decay;
    mm = mm + (1.f - beta1) * (gg - mm);  // lerp, as torch does
    vv = beta2 * vv + (1.f - beta2) * gg * gg;
    pp -= step * (mm / (sqrtf(vv) / sbc2 + eps));
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z);
    upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp;
    m4[i] = mm;
    v4[i] = vv;
    if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float pp = p[i], mm = m[i], vv = v[i];
    upd(pp, g[i], mm, vv);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
    if (zero_grad) g[i] = 0.f;
  }
}

// ---- training noise of the reference's datapipe (src/datasets/base.py:274-289), on the device: per node and output
//      channel c, n ~ N(0, noise_level[c]), zero on nodes whose loss mask is 0 (Dirichlet nodes);
//      node_in[:, c] += n, node_tar[:, c] += (1 - gamma) n.  Philox counter-based stream: (seed, row, offset).
struct NoiseParams {
  float* node_in;
  int Cin;
  float* node_tar;
  int C;
  const float* mask;
  long long rows;
  float level[4];
  float one_minus_gamma;
  unsigned long long seed, offset;
};
__global__ void __launch_bounds__(256) k_inject_noise(const NoiseParams p) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < p.rows; r += (long long)gridDim.x * blockDim.x) {
    if (p.mask[r] == 0.f) continue;
    curandStatePhilox4_32_10_t st;
    curand_init(p.seed, (unsigned long long)r, p.offset, &st);
    const float4 z = curand_normal4(&st);
    const float n[4] = {z.x * p.level[0], z.y * p.level[1], z.z * p.level[2], z.w * p.level[3]};
    for (int c = 0; c < p.C; ++c) {
      p.node_in[r * p.Cin + c] += n[c];
      p.node_tar[r * p.C + c] += p.one_minus_gamma * n[c];
    }
  }
}

static int grid_for(long long n, int per_thread) {
  long long b = (n + 256ll * per_thread - 1) / (256ll * per_thread);
  return (int)std::max<long long>(1, std::min<long long>(b, 148 * 8));
}
}  // namespace bsms

using namespace bsms;

extern "C" int bsms_masked_rmse(const float* pred, const float* tar, const float* mask, int64_t rows, int32_t C,
                                double* acc2_dev, const float* g_loss_dev, float* loss_dev, float* grad_pred, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CHECK_ARG(pred && tar && mask && acc2_dev && rows >= 1 && C >= 1, "bsms_masked_rmse: bad argument");
  BSMS_CUDA(cudaMemsetAsync(acc2_dev, 0, 2 * sizeof(double), st));
  {
    ProfScope ps_(PK_OTHER, st);
    k_rmse_partial<<<grid_for(rows, 4), 256, 0, st>>>(pred, tar, mask, rows, C, acc2_dev);
    BSMS_LAUNCHED();
  }
  {
    ProfScope ps_(PK_OTHER, st);
    k_rmse_finish<<<grad_pred ? grid_for(rows * C, 4) : 1, 256, 0, st>>>(pred, tar, mask, rows, C, acc2_dev, g_loss_dev, loss_dev,
                                                                         grad_pred);
    BSMS_LAUNCHED();
  }
  return BSMS_OK;
}

extern "C" int bsms_clip_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double* state_dev,
                                    float* hyper_dev, double peak_lr, double warmup_steps, double decay_steps, double beta1,
                                    double beta2, double eps, double weight_decay, double max_norm, int32_t zero_grad,
                                    void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state_dev && hyper_dev && n >= 1, "bsms_clip_adamw_step: bad argument");
  BSMS_CHECK_ARG((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                 "bsms_clip_adamw_step: buffers must be 16-byte aligned");
  double* gnorm_sq = state_dev + 1;
  if (max_norm > 0.0) {
    BSMS_CUDA(cudaMemsetAsync(gnorm_sq, 0, sizeof(double), st));
    ProfScope ps_(PK_OTHER, st);
    k_sumsq<<<grid_for(n, 16), 256, 0, st>>>(grads, n, gnorm_sq);
    BSMS_LAUNCHED();
  }
  {
    ProfScope ps_(PK_OTHER, st);
    k_step_scalars<<<1, 1, 0, st>>>(state_dev, max_norm > 0.0 ? gnorm_sq : nullptr, hyper_dev, peak_lr, warmup_steps, decay_steps,
                                    beta1, beta2, max_norm);
    BSMS_LAUNCHED();
  }
  {
    ProfScope ps_(PK_OTHER, st);
    k_clip_adamw<<<grid_for(n, 8), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, hyper_dev, (float)beta1, (float)beta2,
                                                 (float)eps, (float)weight_decay, zero_grad);
    BSMS_LAUNCHED();
  }
  return BSMS_OK;
}

extern "C" int bsms_inject_noise(float* node_in, int32_t Cin, float* node_tar, int32_t C, const float* mask, int64_t rows,
                                 const float* noise_level_host, float noise_gamma, uint64_t seed, uint64_t offset, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CHECK_ARG(node_in && node_tar && mask && noise_level_host && rows >= 1, "bsms_inject_noise: null argument");
  BSMS_CHECK_ARG(C >= 1 && C <= 4 && Cin >= C, "bsms_inject_noise: 1..4 output channels");
  NoiseParams p;
  p.node_in = node_in;
  p.Cin = Cin;
  p.node_tar = node_tar;
  p.C = C;
  p.mask = mask;
  p.rows = rows;
  for (int c = 0; c < 4; ++c) p.level[c] = c < C ? noise_level_host[c] : 0.f;
  p.one_minus_gamma = 1.f - noise_gamma;
  p.seed = seed;
  p.offset = offset;
  ProfScope ps_(PK_OTHER, st);
  k_inject_noise<<<grid_for(rows, 1), 256, 0, st>>>(p);
  BSMS_LAUNCHED();
  return BSMS_OK;
}
