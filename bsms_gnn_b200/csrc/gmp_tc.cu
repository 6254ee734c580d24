// GMP block on the tensor-core path (BSMS_MODE_BF16 / BSMS_MODE_FP16X3): orchestration of the fused
// edge kernels (edge_chain*.cu) and the node-level tcgen05 GEMMs (node_gemm.cu).
// Reference: src/ops/basic.py:48-98.
#include "chain.cuh"

namespace bsms {

// kernels / launchers defined in the other translation units
int edge_chain_forward(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* PsPd, const float* pos,
                       int pos_batched, int B, int P, int mode, uint8_t* wpack, float* aggr, float* dbg, int dbg_stage,
                       cudaStream_t st, bool prepacked, uint8_t* bpack, float* eout = nullptr);
int edge_chain_backward(const bsms_level_plan* pl, const bsms_gmp_weights* w, const bsms_gmp_grads* gr, const float* PsPd,
                        const float* pos, int pos_batched, int B, int P, uint8_t* wpack, const float* g_aggr, int ld_g,
                        float* gPsPd, cudaStream_t st, bool prepacked, float* g0_rows = nullptr, float* part = nullptr);
size_t gmp_pack_stride(int mode);
int gmp_pack_blocks(const PackList& pl, int mode, uint8_t* out, cudaStream_t st);
int edge_chain_pack_bias(const bsms_gmp_weights* w, uint8_t* bpack, cudaStream_t st);
int lin_tc(int mode, const float* X0, int ldx0, const float* X1, int ldx1, int KB, int NB, const uint8_t* const* blocks,
           int b_mn, const float* bias, int relu, const float* mask, int ldmask, int accum, float* Y, int ldy,
           long long rows, int kind, cudaStream_t st);
int lin_tc2(const float* X0, int ldx0, const float* X1, int ldx1, int KB, int NB, const uint8_t* const* blocks, int b_mn,
            const float* bias, int relu, const float* mask, int ldmask, const float* add0, int ldadd0, const float* add1,
            int ldadd1, float* Y0, int ldy0, float* Y1, int ldy1, float* ln_out, const float* res0, const float* res1,
            long long rows, int kind, cudaStream_t st);
int wgrad_tc_batch(const WgradParams* probs, int nprob, cudaStream_t st, float* part = nullptr);
int node_chain_backward(const float* Yn, const float* g_out, const float* N1, const float* N2, const float* N3, int img,
                        const uint8_t* wpack_v2, float* G4, float* const* gW, float* const* gb, long long rows,
                        cudaStream_t st, float* part = nullptr);
int launch_edge_grad_segsum(const float* g0_rows, const bsms_level_plan* pl, float* gPsPd, int B, cudaStream_t st);  // gmp.cu
int node_chain_forward(const float* N1, const uint8_t* wpack_v2, const float* c2, const float* c3, const float* c4,
                       uint8_t* img2, uint8_t* img3, float* Yn, const float* x, const float* skip, float* out,
                       long long rows, cudaStream_t st);

// The fused node chains keep N2 / N3 as bf16 operand-tile images inside the fp32-sized buffers of the node
// tensors: ceil(Rn/128) x 32 KB fits into Rn x 512 B from 64 rows on; smaller problems (the coarsest levels of
// small meshes) run the layer-by-layer kernels with fp32 activations.
static inline bool node_images(long long Rn) { return Rn >= 64; }
WgradParams wgrad_problem(const float* G, int ldg, const float* X, int ldx, float* dW, int ldo, float* db, long long rows);
int launch_ln_residual(const float* Yn, const float* x, const float* skip, float* out, long long rows, cudaStream_t st);

#define TC_TRY(expr)              \
  do {                            \
    int _rc = (expr);             \
    if (_rc != BSMS_OK) return _rc; \
  } while (0)

// packed block indices of one GMP
enum { BW2 = 0, BW3, BW4, BW1S, BW1D, BV1A, BV1B, BV2, BV3, BV4, NBLOCKS };

// bias of the fused projection GEMM: [0 (Ps half) | b1 (Pd half)] — b1 rides with Pd into the edge kernels
__global__ void k_bias_sd(const float* __restrict__ b1, float* __restrict__ out) {
  const int i = threadIdx.x;
  out[i] = i < 128 ? 0.f : b1[i - 128];
}

// scratch layout behind the packed weight blocks
constexpr size_t kPackBytes = (size_t)10 * 2 * kWBlk;   // NBLOCKS blocks, hi+lo
constexpr size_t kBiasPackOff = kPackBytes;             // three 16 KB bias blocks of the edge chain
constexpr size_t kBiasSdOff = kPackBytes + 3 * 16384;   // 256 floats
constexpr size_t kScratchBytes = kBiasSdOff + 1024;

static int pack_all(const bsms_gmp_weights* w, int P, int mode, uint8_t* wpack, cudaStream_t st) {
  {
    ProfScope ps_(PK_OTHER, st);
    k_bias_sd<<<1, 256, 0, st>>>(w->b_edge[0], reinterpret_cast<float*>(wpack + kBiasSdOff));
    BSMS_LAUNCHED();
  }
  const int ldw1 = 2 * kD + P + 1;
  PackList pl;
  pl.n = NBLOCKS;
  pl.w[BW2] = w->w_edge[1]; pl.ld[BW2] = kD;
  pl.w[BW3] = w->w_edge[2]; pl.ld[BW3] = kD;
  pl.w[BW4] = w->w_edge[3]; pl.ld[BW4] = kD;
  pl.w[BW1S] = w->w_edge[0] + (P + 1); pl.ld[BW1S] = ldw1;
  pl.w[BW1D] = w->w_edge[0] + (P + 1 + kD); pl.ld[BW1D] = ldw1;
  pl.w[BV1A] = w->w_node[0]; pl.ld[BV1A] = 2 * kD;
  pl.w[BV1B] = w->w_node[0] + kD; pl.ld[BV1B] = 2 * kD;
  pl.w[BV2] = w->w_node[1]; pl.ld[BV2] = kD;
  pl.w[BV3] = w->w_node[2]; pl.ld[BV3] = kD;
  pl.w[BV4] = w->w_node[3]; pl.ld[BV4] = kD;
  TC_TRY(gmp_pack_blocks(pl, mode, wpack, st));
  if (mode == BSMS_MODE_BF16) TC_TRY(edge_chain_pack_bias(w, wpack + kBiasPackOff, st));  // b2..b4 as one MMA operand block
  return BSMS_OK;
}

struct NodeBufs {
  float *PsPd, *aggr, *N1, *N2, *N3, *Yn;
};
// same layout as gmp.cu's carve() so `saved` is interchangeable between the paths
static NodeBufs carve_nodes(Arena& a, long long Rn) {
  NodeBufs n;
  n.PsPd = a.take<float>(Rn * 256);
  n.aggr = a.take<float>(Rn * kD);
  n.N1 = a.take<float>(Rn * kD);
  n.N2 = a.take<float>(Rn * kD);
  n.N3 = a.take<float>(Rn * kD);
  n.Yn = a.take<float>(Rn * kD);
  return n;
}

// ---- deterministic option ------------------------------------------------------------------------------------
struct DetSegs {
  DetSeg s[12];
  const float* part;
  int stride;
};
// dst[row, col] += sum over the CTAs' partial blocks and their per-warp repetitions, in ONE fixed order: 64 elements per
// CTA, eight interleaved chains per element (thread group j takes blocks j, j+8, ... into one accumulator and blocks
// j+4, j+12, ... into a second one), combined in a fixed sequence — every run adds the same numbers in the same order
__global__ void __launch_bounds__(256) k_det_reduce(const DetSegs a) {
  __shared__ float s_part[4][64];
  const DetSeg sg = a.s[blockIdx.y];
  const int e = threadIdx.x & 63, j = threadIdx.x >> 6;
  const int idx = blockIdx.x * 64 + e;
  const bool live = idx < sg.rows * sg.cols;
  float acc0 = 0.f, acc1 = 0.f;
  int row = 0, col = 0;
  if (live) {
    row = idx / sg.cols;
    col = idx - row * sg.cols;
    const float* src = a.part + (size_t)sg.part0 * a.stride + sg.src_off + row * sg.src_cols + col;
    const size_t rep_stride = (size_t)sg.rows * sg.src_cols;
    for (int c = j; c < sg.nparts; c += 8) {
      const float* p0 = src + (size_t)c * a.stride;
      for (int rep = 0; rep < sg.reps; ++rep) acc0 += p0[rep * rep_stride];
      if (c + 4 < sg.nparts) {
        const float* p1 = p0 + (size_t)4 * a.stride;
        for (int rep = 0; rep < sg.reps; ++rep) acc1 += p1[rep * rep_stride];
      }
    }
  }
  s_part[j][e] = acc0 + acc1;
  __syncthreads();
  if (live && j == 0) sg.dst[(size_t)row * sg.ldd + col] += ((s_part[0][e] + s_part[1][e]) + s_part[2][e]) + s_part[3][e];
}

int det_reduce(const float* part, int stride, const DetSeg* segs, int nseg, cudaStream_t st) {
  if (nseg < 1 || nseg > 12) {
    set_error("det_reduce: bad segment count");
    return BSMS_EINVAL;
  }
  DetSegs a;
  int most = 1;
  for (int i = 0; i < nseg; ++i) {
    a.s[i] = segs[i];
    most = std::max(most, segs[i].rows * segs[i].cols);
  }
  a.part = part;
  a.stride = stride;
  ProfScope ps_(PK_WGRAD, st);
  k_det_reduce<<<dim3(ceil_div(most, 64), nseg), 256, 0, st>>>(a);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

static int det_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
// one partial block per CTA of the widest flush (the fused edge backward); every flushing kernel launches <= #SMs CTAs
size_t det_part_bytes() { return (size_t)det_sms() * kDetEdgeBwdStride * sizeof(float); }

// aggr[b, n] = sum of the normalised edge rows of node n's in-edges, in CSR (dst-sorted) order: warp per node row
__global__ void __launch_bounds__(256) k_segsum_rows(const float* __restrict__ rows, const int32_t* __restrict__ rowptr_d,
                                                     float* __restrict__ aggr, int B, int N, int E) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (long long)B * N) return;
  const int b = (int)(gw / N), n = (int)(gw - (long long)b * N);
  const float* base = rows + (size_t)b * E * kD + 4 * lane;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = rowptr_d[n]; k < rowptr_d[n + 1]; ++k) {
    const float4 v = ld4(base + (size_t)k * kD);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  st4(aggr + gw * kD + 4 * lane, acc);
}

// the fused bf16 edge stage into aggr: red.add reduction (default) or rows + ordered segment sum (deterministic)
static int edge_stage_bf16(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* PsPd, const float* pos,
                           int pos_batched, int B, int P, uint8_t* wpack, float* aggr, float* det_rows, cudaStream_t st) {
  const long long Rn = (long long)B * pl->n_nodes, Re = (long long)B * pl->n_edges;
  if (!det_rows || Re == 0) {
    BSMS_CUDA(cudaMemsetAsync(aggr, 0, (size_t)Rn * kD * sizeof(float), st));
    return edge_chain_forward(pl, w, PsPd, pos, pos_batched, B, P, BSMS_MODE_BF16, wpack, aggr, nullptr, -1, st, true,
                              wpack + kBiasPackOff);
  }
  TC_TRY(edge_chain_forward(pl, w, PsPd, pos, pos_batched, B, P, BSMS_MODE_BF16, wpack, aggr, nullptr, -1, st, true,
                            wpack + kBiasPackOff, det_rows));
  ProfScope ps_(PK_LN_SEGSUM, st);
  k_segsum_rows<<<ceil_div(Rn * 32, 256), 256, 0, st>>>(det_rows, pl->rowptr_d, aggr, B, pl->n_nodes, pl->n_edges);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

// out != nullptr: the last layer also writes out = LN(Yn) + x (+ skip) (fused in the bf16 mode)
static int forward_nodes(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos,
                         int pos_batched, int B, int P, int mode, const NodeBufs& n, uint8_t* wpack, const float* skip,
                         float* out, cudaStream_t st, float* det_rows = nullptr) {
  const long long Rn = (long long)B * pl->n_nodes;
  const size_t bs = gmp_pack_stride(mode);
  auto blk = [&](int i) { return (const uint8_t*)(wpack + (size_t)i * bs); };
  if (mode == BSMS_MODE_BF16) {
    const int K = PK_NODE_FWD_GEMM;
    {
      const uint8_t* b[2] = {blk(BW1S), blk(BW1D)};  // Ps | Pd = x [W1s ; W1d]^T
      TC_TRY(lin_tc2(x, kD, nullptr, 0, 1, 2, b, 0, reinterpret_cast<const float*>(wpack + kBiasSdOff), 0, nullptr, 0, nullptr,
                     0, nullptr, 0, n.PsPd, 256, n.PsPd + 128, 256, nullptr, nullptr, nullptr, Rn, K, st));
    }
    TC_TRY(edge_stage_bf16(pl, w, n.PsPd, pos, pos_batched, B, P, wpack, n.aggr, det_rows, st));
    {
      const uint8_t* b[2] = {blk(BV1A), blk(BV1B)};  // N1 = relu([x | aggr] V1^T + c1)
      TC_TRY(lin_tc2(x, kD, n.aggr, kD, 2, 1, b, 0, w->b_node[0], 1, nullptr, 0, nullptr, 0, nullptr, 0, n.N1, kD, nullptr, 0,
                     nullptr, nullptr, nullptr, Rn, K, st));
    }
    if (node_images(Rn)) {
      // layers 1..3, LayerNorm and the residual(s) in one kernel; N2 / N3 are kept as operand-tile images
      TC_TRY(node_chain_forward(n.N1, blk(BV2), w->b_node[1], w->b_node[2], w->b_node[3], reinterpret_cast<uint8_t*>(n.N2),
                                reinterpret_cast<uint8_t*>(n.N3), n.Yn, x, skip, out, Rn, st));
      return BSMS_OK;
    }
    const float* in[3] = {n.N1, n.N2, n.N3};
    float* o[3] = {n.N2, n.N3, n.Yn};
    const int bi[3] = {BV2, BV3, BV4};
    for (int l = 0; l < 3; ++l) {
      const uint8_t* b[1] = {blk(bi[l])};
      const bool last = l == 2;
      TC_TRY(lin_tc2(in[l], kD, nullptr, 0, 1, 1, b, 0, w->b_node[l + 1], last ? 0 : 1, nullptr, 0, nullptr, 0, nullptr, 0, o[l],
                     kD, nullptr, 0, last ? out : nullptr, x, skip, Rn, K, st));
    }
    return BSMS_OK;
  }
  {
    const uint8_t* b[2] = {blk(BW1S), blk(BW1D)};  // Ps | Pd = x [W1s ; W1d]^T
    TC_TRY(lin_tc(mode, x, kD, nullptr, 0, 1, 2, b, 0, reinterpret_cast<const float*>(wpack + kBiasSdOff), 0, nullptr, 0, 0,
                  n.PsPd, 256, Rn, PK_NODE_FWD_GEMM, st));
  }
  BSMS_CUDA(cudaMemsetAsync(n.aggr, 0, (size_t)Rn * kD * sizeof(float), st));
  TC_TRY(edge_chain_forward(pl, w, n.PsPd, pos, pos_batched, B, P, mode, wpack, n.aggr, nullptr, -1, st, true,
                            wpack + kBiasPackOff));
  {
    const uint8_t* b[2] = {blk(BV1A), blk(BV1B)};  // N1 = relu([x | aggr] V1^T + c1)
    TC_TRY(lin_tc(mode, x, kD, n.aggr, kD, 2, 1, b, 0, w->b_node[0], 1, nullptr, 0, 0, n.N1, kD, Rn, PK_NODE_FWD_GEMM, st));
  }
  {
    const uint8_t* b[1] = {blk(BV2)};
    TC_TRY(lin_tc(mode, n.N1, kD, nullptr, 0, 1, 1, b, 0, w->b_node[1], 1, nullptr, 0, 0, n.N2, kD, Rn, PK_NODE_FWD_GEMM, st));
  }
  {
    const uint8_t* b[1] = {blk(BV3)};
    TC_TRY(lin_tc(mode, n.N2, kD, nullptr, 0, 1, 1, b, 0, w->b_node[2], 1, nullptr, 0, 0, n.N3, kD, Rn, PK_NODE_FWD_GEMM, st));
  }
  {
    const uint8_t* b[1] = {blk(BV4)};
    TC_TRY(lin_tc(mode, n.N3, kD, nullptr, 0, 1, 1, b, 0, w->b_node[3], 0, nullptr, 0, 0, n.Yn, kD, Rn, PK_NODE_FWD_GEMM, st));
  }
  if (out) return launch_ln_residual(n.Yn, x, skip, out, Rn, st);
  return BSMS_OK;
}

size_t gmp_packed_bytes() { return kScratchBytes; }
int gmp_pack_tc(const bsms_gmp_weights* w, int P, int mode, uint8_t* packed, cudaStream_t st) { return pack_all(w, P, mode, packed, st); }

// `packed` (may be null): the images made by bsms_gmp_pack for these weights in this mode — nothing is packed here
int gmp_forward_tc(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos, int pos_batched,
                   const float* skip, float* out, float* saved, int B, int P, int mode, void* ws, size_t ws_bytes,
                   cudaStream_t st, const uint8_t* packed) {
  const long long Rn = (long long)B * pl->n_nodes;
  Arena ar(ws, ws_bytes);
  Arena sv(saved, (size_t)-1);
  NodeBufs n = carve_nodes(saved ? sv : ar, Rn);
  // with `saved` the packed weight images are kept next to the node tensors so that backward does not re-pack
  uint8_t* wpack = saved ? sv.take<uint8_t>(kScratchBytes) : ar.take<uint8_t>(kScratchBytes);
  float* det_rows = nullptr;
  if (det_enabled() && mode == BSMS_MODE_BF16) det_rows = ar.take<float>(std::max<long long>((long long)B * pl->n_edges, 1) * kD);
  if (!ar.ok()) {
    set_error("bsms_gmp_forward: workspace too small for the tensor-core path");
    return BSMS_EWORKSPACE;
  }
  if (packed && !saved)
    wpack = const_cast<uint8_t*>(packed);
  else if (packed)
    BSMS_CUDA(cudaMemcpyAsync(wpack, packed, kScratchBytes, cudaMemcpyDeviceToDevice, st));
  else
    TC_TRY(pack_all(w, P, mode, wpack, st));
  return forward_nodes(pl, w, x, pos, pos_batched, B, P, mode, n, wpack, skip, out, st, det_rows);
}

// bf16 backward: node MLP backward on tcgen05 GEMMs, fused edge backward, node-level layer-0 gradients
int gmp_backward_tc(const bsms_level_plan* pl, const bsms_gmp_weights* w, const float* x, const float* pos,
                    int pos_batched, const float* saved, const float* g_out, float* g_x, const bsms_gmp_grads* gr, int B,
                    int P, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int mode = BSMS_MODE_BF16;
  const long long Rn = (long long)B * pl->n_nodes, Re = (long long)B * pl->n_edges;
  const int ldw1 = 2 * kD + P + 1;
  Arena ar(ws, ws_bytes);
  Arena sv(const_cast<float*>(saved), (size_t)-1);
  NodeBufs n = carve_nodes(saved ? sv : ar, Rn);
  uint8_t* saved_pack = saved ? sv.take<uint8_t>(kScratchBytes) : nullptr;  // packed by forward (same weights)
  ar.take<float>(Rn * kD * 3);  // (three gradient buffers of the unfused path; kept so the workspace formula is unchanged)
  float* G4 = ar.take<float>(Rn * kD);
  float* gcat = ar.take<float>(Rn * 256);  // only the first Rn*128 floats are used (g_aggr)
  float* gPsPd = ar.take<float>(Rn * 256);
  uint8_t* wpack = ar.take<uint8_t>(kScratchBytes);
  // deterministic option: edge rows (forward recompute / edge-input gradient) and the per-CTA partial-sum block
  float* det_rows = nullptr;
  float* det_part = nullptr;
  if (det_enabled()) {
    det_rows = ar.take<float>(std::max<long long>(Re, 1) * kD);
    det_part = reinterpret_cast<float*>(ar.take<uint8_t>(det_part_bytes()));
  }
  if (!ar.ok()) {
    set_error("bsms_gmp_backward: workspace too small for the tensor-core path");
    return BSMS_EWORKSPACE;
  }
  const size_t bs = gmp_pack_stride(mode);
  if (saved_pack)
    wpack = saved_pack;
  else
    TC_TRY(pack_all(w, P, mode, wpack, st));
  auto blk = [&](int i) { return (const uint8_t*)(wpack + (size_t)i * bs); };
  if (!saved) TC_TRY(forward_nodes(pl, w, x, pos, pos_batched, B, P, mode, n, wpack, nullptr, nullptr, st, det_rows));
  // ---- node MLP backward, layers 3..1: LayerNorm backward, data- and weight-gradient GEMMs in one fused kernel
  {
    float* gWn[3] = {gr->w_node[1], gr->w_node[2], gr->w_node[3]};
    float* gbn[3] = {gr->b_node[1], gr->b_node[2], gr->b_node[3]};
    TC_TRY(node_chain_backward(n.Yn, g_out, n.N1, n.N2, n.N3, node_images(Rn) ? 1 : 0, blk(BV2), G4, gWn, gbn, Rn, st, det_part));
  }
  {
    WgradParams pr[2] = {wgrad_problem(G4, kD, x, kD, gr->w_node[0], 2 * kD, gr->b_node[0], Rn),  // layer 0: input [x | aggr]
                         wgrad_problem(G4, kD, n.aggr, kD, gr->w_node[0] + kD, 2 * kD, nullptr, Rn)};
    TC_TRY(wgrad_tc_batch(pr, 2, st, det_part));
  }
  {
    // [g_x | g_aggr] = G4 V1: the x half leaves as g_x = g_out + ., the aggr half feeds the edge backward
    const uint8_t* b[2] = {blk(BV1A), blk(BV1B)};
    TC_TRY(lin_tc2(G4, kD, nullptr, 0, 1, 2, b, 1, nullptr, 0, nullptr, 0, g_out, kD, nullptr, 0, g_x, kD, gcat, kD, nullptr,
                   nullptr, nullptr, Rn, PK_DGRAD, st));
  }
  // ---- edge stage backward (fused) and the node-level gradients of the first edge layer
  if (Re > 0) {
    if (det_rows) {
      // the edge-input gradient leaves the fused kernel as rows; gPs | gPd are CSR-ordered segment sums of them
      TC_TRY(edge_chain_backward(pl, w, gr, n.PsPd, pos, pos_batched, B, P, wpack, gcat, kD, gPsPd, st, true, det_rows, det_part));
      TC_TRY(launch_edge_grad_segsum(det_rows, pl, gPsPd, B, st));
    } else {
      BSMS_CUDA(cudaMemsetAsync(gPsPd, 0, (size_t)Rn * 256 * sizeof(float), st));
      TC_TRY(edge_chain_backward(pl, w, gr, n.PsPd, pos, pos_batched, B, P, wpack, gcat, kD, gPsPd, st, true));
    }
    WgradParams pr[2] = {wgrad_problem(gPsPd, 256, x, kD, gr->w_edge[0] + (P + 1), ldw1, nullptr, Rn),
                         wgrad_problem(gPsPd + 128, 256, x, kD, gr->w_edge[0] + (P + 1 + kD), ldw1, nullptr, Rn)};
    TC_TRY(wgrad_tc_batch(pr, 2, st, det_part));
    const uint8_t* b[2] = {blk(BW1S), blk(BW1D)};  // g_x += gPs W1s + gPd W1d
    TC_TRY(lin_tc2(gPsPd, 256, gPsPd + 128, 256, 2, 1, b, 1, nullptr, 0, nullptr, 0, g_x, kD, nullptr, 0, g_x, kD, nullptr, 0,
                   nullptr, nullptr, nullptr, Rn, PK_DGRAD, st));
  }
  return BSMS_OK;
}

}  // namespace bsms
