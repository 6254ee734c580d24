// Level-plan construction: int64 [2,E] edge list -> int32 dst-sorted and src-sorted CSR views.
// One-time per mesh level (reference keeps raw edge lists and re-derives everything per op,
// src/ops/basic.py:66,130-137).  Uses CUB radix sort (stable) + scan from the CUDA toolkit.
#include <cub/cub.cuh>
#include <stdarg.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace bsms {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- profiling registry
struct ProfRec {
  int kind;
  cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
bool prof_enabled() { return g_prof_on; }
void prof_begin(int kind, cudaStream_t st) {
  ProfRec r;
  r.kind = kind;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) { cudaEventRecord(g_prof.back().b, st); }

// keys[0..E) = g[which], vals = iota; status[0] |= out-of-range flag
__global__ void k_plan_keys(const int64_t* __restrict__ g, int64_t E, int64_t N, int32_t* __restrict__ ksrc,
                            int32_t* __restrict__ kdst, int32_t* __restrict__ iota, int32_t* __restrict__ status) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = g[e], d = g[E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) {
    atomicOr(&status[0], 1);
    s = 0;
    d = 0;
  }
  ksrc[e] = (int32_t)s;
  kdst[e] = (int32_t)d;
  iota[e] = (int32_t)e;
  // status[1] = max sender index: the reference's degree() sizes itself by it (src/utils/basic.py:305-307)
  const int32_t wmax = __reduce_max_sync(__activemask(), (int32_t)s);
  if ((threadIdx.x & 31) == 0 || e == 0) atomicMax(&status[1], wmax);
}

// Content fingerprint of up to 32 buffers of 8-byte words in ONE launch: out[k] = sum_i mix(word_i, i) mod 2^64
// (splitmix64 finaliser; position-dependent, order of accumulation irrelevant).  blockIdx.y = buffer.
struct FingerprintArgs {
  const unsigned long long* ptr[32];
  long long nwords[32];
};
__global__ void k_fingerprint(FingerprintArgs a, unsigned long long* __restrict__ out) {
  const int k = blockIdx.y;
  const unsigned long long* p = a.ptr[k];
  const long long n = a.nwords[k];
  unsigned long long acc = 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = p[i] + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    acc += z ^ (z >> 31);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out + k, acc);
}

// gather the other endpoint through the sorted permutation; histogram of the sort key
__global__ void k_plan_gather(const int32_t* __restrict__ perm, const int32_t* __restrict__ other, int64_t E,
                              int32_t* __restrict__ other_sorted, const int32_t* __restrict__ key_sorted,
                              int32_t* __restrict__ counts) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  other_sorted[k] = other[perm[k]];
  atomicAdd(&counts[key_sorted[k]], 1);
}

// inverse permutation: pos_d[perm_d[k]] = k, then s2d[k'] = pos_d[perm_s[k']]
__global__ void k_plan_invert(const int32_t* __restrict__ perm_d, int64_t E, int32_t* __restrict__ pos_d) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < E) pos_d[perm_d[k]] = (int32_t)k;
}
__global__ void k_plan_s2d(const int32_t* __restrict__ perm_s, const int32_t* __restrict__ pos_d, int64_t E,
                           int32_t* __restrict__ s2d) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < E) s2d[k] = pos_d[perm_s[k]];
}
}  // namespace bsms

using namespace bsms;

extern "C" const char* bsms_last_error(void) { return g_err; }
extern "C" int bsms_version(void) { return 100; }
extern "C" int64_t bsms_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int bsms_prof_enable(int on) {
  g_prof_on = on != 0;
  return BSMS_OK;
}
// Sums the recorded kernel durations per kind (ms) and launch counts, then clears the records.
extern "C" int bsms_prof_collect(double* ms_by_kind, int64_t* launches_by_kind, int n_kinds) {
  BSMS_CHECK_ARG(ms_by_kind && launches_by_kind && n_kinds >= PK_COUNT, "bsms_prof_collect: need %d kinds", PK_COUNT);
  for (int k = 0; k < n_kinds; ++k) {
    ms_by_kind[k] = 0.0;
    launches_by_kind[k] = 0;
  }
  BSMS_CUDA(cudaDeviceSynchronize());
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    ms_by_kind[r.kind] += ms;
    launches_by_kind[r.kind] += 1;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return BSMS_OK;
}

// out_dev[k] (zeroed here) = fingerprint of buffer k; ptrs / nbytes are HOST arrays of n <= 32 entries, sizes
// multiples of 8.  Asynchronous on `stream`: the host reads out_dev after synchronising.
extern "C" int bsms_fingerprint(const void* const* ptrs, const int64_t* nbytes, int32_t n, uint64_t* out_dev, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  BSMS_CHECK_ARG(ptrs && nbytes && out_dev && n >= 1 && n <= 32, "bsms_fingerprint: 1..32 buffers");
  FingerprintArgs a;
  long long most = 0;
  for (int k = 0; k < 32; ++k) {
    a.ptr[k] = nullptr;
    a.nwords[k] = 0;
  }
  for (int k = 0; k < n; ++k) {
    BSMS_CHECK_ARG(nbytes[k] >= 0 && nbytes[k] % 8 == 0 && (nbytes[k] == 0 || ptrs[k]), "bsms_fingerprint: buffer %d", k);
    a.ptr[k] = (const unsigned long long*)ptrs[k];
    a.nwords[k] = nbytes[k] / 8;
    most = std::max(most, a.nwords[k]);
  }
  BSMS_CUDA(cudaMemsetAsync(out_dev, 0, (size_t)n * sizeof(uint64_t), st));
  const int bx = (int)std::min<long long>(std::max<long long>(ceil_div(most, 256 * 8), 1), 296);
  k_fingerprint<<<dim3(bx, n), 256, 0, st>>>(a, (unsigned long long*)out_dev);
  BSMS_LAUNCHED();
  return BSMS_OK;
}

extern "C" int bsms_device_info(int64_t* out) {
  BSMS_CHECK_ARG(out != nullptr, "bsms_device_info: null output");
  int dev = 0;
  BSMS_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  BSMS_CUDA(cudaGetDeviceProperties(&p, dev));
  out[0] = p.multiProcessorCount;
  out[1] = p.l2CacheSize;
  out[2] = (int64_t)p.sharedMemPerBlockOptin;
  out[3] = p.major * 10 + p.minor;
  return BSMS_OK;
}

static size_t sort_temp_bytes(int64_t E) {
  size_t t = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)E);
  return t;
}
static size_t scan_temp_bytes(int64_t N) {
  size_t t = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, t, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(N + 1));
  return t;
}

extern "C" size_t bsms_plan_workspace_bytes(int64_t E, int64_t N) {
  size_t e = align_up((size_t)(E > 0 ? E : 1) * 4, 256);
  size_t n = align_up((size_t)(N + 1) * 4, 256);
  return 4 * e + n + align_up(sort_temp_bytes(E > 0 ? E : 1), 256) + align_up(scan_temp_bytes(N), 256) + 1024;
}

extern "C" int bsms_plan_build(const int64_t* g, int64_t E, int64_t N, const bsms_level_plan* out, int32_t* status,
                               void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  BSMS_CHECK_ARG(out && status && ws, "bsms_plan_build: null argument");
  BSMS_CHECK_ARG(N >= 1 && N < (1ll << 31) && E >= 0 && E < (1ll << 31), "bsms_plan_build: sizes out of int32 range");
  BSMS_CHECK_ARG(ws_bytes >= bsms_plan_workspace_bytes(E, N), "bsms_plan_build: workspace too small");
  int32_t* rp_d = (int32_t*)out->rowptr_d;
  int32_t* rp_s = (int32_t*)out->rowptr_s;
  BSMS_CUDA(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
  if (E == 0) {
    BSMS_CUDA(cudaMemsetAsync(rp_d, 0, (N + 1) * 4, st));
    BSMS_CUDA(cudaMemsetAsync(rp_s, 0, (N + 1) * 4, st));
    return BSMS_OK;
  }
  BSMS_CHECK_ARG(g != nullptr, "bsms_plan_build: null edge list");
  Arena a(ws, ws_bytes);
  int32_t* ksrc = a.take<int32_t>(E);
  int32_t* kdst = a.take<int32_t>(E);
  int32_t* iota = a.take<int32_t>(E);
  int32_t* pos_d = a.take<int32_t>(E);
  int32_t* counts = a.take<int32_t>(N + 1);
  size_t tb_sort = sort_temp_bytes(E), tb_scan = scan_temp_bytes(N);
  void* t_sort = a.take<char>(tb_sort);
  void* t_scan = a.take<char>(tb_scan);
  const int T = 256;
  int nb = ceil_div(E, T);
  k_plan_keys<<<nb, T, 0, st>>>(g, E, N, ksrc, kdst, iota, status);
  BSMS_LAUNCHED();
  // end_bit: only the bits N needs
  int bits = 1;
  while ((1ll << bits) < N) ++bits;
  // --- dst-sorted view
  BSMS_CUDA(cub::DeviceRadixSort::SortPairs(t_sort, tb_sort, kdst, (int32_t*)out->dst_d, iota, (int32_t*)out->perm_d,
                                            (int)E, 0, bits, st));
  BSMS_CUDA(cudaMemsetAsync(counts, 0, (N + 1) * 4, st));
  k_plan_gather<<<nb, T, 0, st>>>(out->perm_d, ksrc, E, (int32_t*)out->src_d, out->dst_d, counts);
  BSMS_LAUNCHED();
  BSMS_CUDA(cub::DeviceScan::ExclusiveSum(t_scan, tb_scan, counts, rp_d, (int)(N + 1), st));
  // --- src-sorted view (perm_s is scratch: reuse iota's slot after sorting into pos_d... keep simple)
  int32_t* perm_s = pos_d;  // temporarily holds perm_s
  BSMS_CUDA(cub::DeviceRadixSort::SortPairs(t_sort, tb_sort, ksrc, (int32_t*)out->src_s, iota, perm_s, (int)E, 0, bits,
                                            st));
  BSMS_CUDA(cudaMemsetAsync(counts, 0, (N + 1) * 4, st));
  k_plan_gather<<<nb, T, 0, st>>>(perm_s, kdst, E, (int32_t*)out->dst_s, out->src_s, counts);
  BSMS_LAUNCHED();
  BSMS_CUDA(cub::DeviceScan::ExclusiveSum(t_scan, tb_scan, counts, rp_s, (int)(N + 1), st));
  // --- cross map: ksrc is free now -> holds inverse of perm_d
  int32_t* inv_d = ksrc;
  k_plan_invert<<<nb, T, 0, st>>>(out->perm_d, E, inv_d);
  BSMS_LAUNCHED();
  k_plan_s2d<<<nb, T, 0, st>>>(perm_s, inv_d, E, (int32_t*)out->s2d);
  BSMS_LAUNCHED();
  int32_t h_status[4];
  BSMS_CUDA(cudaMemcpyAsync(h_status, status, sizeof(h_status), cudaMemcpyDeviceToHost, st));
  BSMS_CUDA(cudaStreamSynchronize(st));
  if (h_status[0] != 0) {
    set_error("bsms_plan_build: edge index out of range [0, %lld)", (long long)N);
    return BSMS_EINDEX;
  }
  return BSMS_OK;
}
