// K6: halo exchange of the node-partitioned processor as ONE kernel per exchange over peer-mapped memory
// (the reference has no partitioning).  Every rank owns an arena allocated with cudaMalloc
// and exported with CUDA IPC; its peers on the same NVLink/NVSwitch box map it and the kernel below stores
// ghost rows straight into the consumer's buffer — no packing buffer, no NCCL group, no unpack kernel:
//
//   forward  ("refresh ghosts"):   out[0:n_own] = x_own (local copy);  for every peer q: the rows q needs from me
//                                  -> q's out buffer, at the ghost rows that belong to me;  signal q;  wait until
//                                  every owner of my ghosts has signalled.
//   backward (adjoint):            g_own = g[0:n_own];  my ghost-row gradients -> the owners' `back` regions;
//                                  signal; wait;  g_own[send_idx[i]] += back[i]  (red.add: a row sent to two peers
//                                  receives two contributions).
//
// Signals are per-(call site, peer) 32-bit epochs written with st.release.sys after a system-scope fence and
// read with ld.acquire.sys; the epoch of a launch is read from device memory (ctrl[0] + 1) so that the kernel
// can be replayed from a CUDA graph.  All CTAs of a launch are co-resident (grid <= 2 x SMs, no shared memory),
// which makes the in-kernel waits and the two arrive counters safe.
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace bsms {

constexpr int kMaxWorld = 8;

struct HaloArgs {
  int world, rank, C4, backward;
  long long n_own, n_ghost, n_send;
  const float4* src;
  float4* dst;
  const long long* send_idx;
  int send_off[kMaxWorld + 1];
  int recv_off[kMaxWorld + 1];
  float4* peer_dst[kMaxWorld];
  const float4* back;
  unsigned* my_flags;
  unsigned* peer_flag[kMaxWorld];
  unsigned* ctrl;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_volatile(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) k_halo(const HaloArgs a) {
  const unsigned e = ld_volatile(a.ctrl) + 1u;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  const int C4 = a.C4;
  // ---- phase 1: remote pushes first (they have the longest way to go), then the local copy of the owned rows
  if (!a.backward) {
    for (int q = 0; q < a.world; ++q) {
      const long long cnt = (long long)(a.send_off[q + 1] - a.send_off[q]) * C4;
      const long long* idx = a.send_idx + a.send_off[q];
      float4* dq = a.peer_dst[q];
      for (long long i = gtid; i < cnt; i += gstride) {
        const long long r = i / C4;
        const int c = (int)(i - r * C4);
        dq[i] = a.src[idx[r] * C4 + c];
      }
    }
  } else {
    for (int q = 0; q < a.world; ++q) {
      const long long cnt = (long long)(a.recv_off[q + 1] - a.recv_off[q]) * C4;
      const float4* sq = a.src + (a.n_own + a.recv_off[q]) * C4;
      float4* dq = a.peer_dst[q];
      for (long long i = gtid; i < cnt; i += gstride) dq[i] = sq[i];
    }
  }
  for (long long i = gtid; i < a.n_own * C4; i += gstride) a.dst[i] = a.src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(a.ctrl + 1, 1u);
    if (t == gridDim.x - 1) {  // last CTA of this rank: every push above is visible system-wide
      a.ctrl[1] = 0u;
      __threadfence_system();
      for (int q = 0; q < a.world; ++q) {
        const int out_cnt = a.backward ? a.recv_off[q + 1] - a.recv_off[q] : a.send_off[q + 1] - a.send_off[q];
        if (q != a.rank && out_cnt > 0) st_release_sys(a.peer_flag[q], e);
      }
      st_release_sys(a.ctrl + 2, e);  // local: phase 1 is complete on this rank
    }
  }
  // ---- phase 2: wait for every peer that sends to me
  if (threadIdx.x < a.world) {
    const int q = threadIdx.x;
    const int in_cnt = a.backward ? a.send_off[q + 1] - a.send_off[q] : a.recv_off[q + 1] - a.recv_off[q];
    if (q != a.rank && in_cnt > 0) {
      while ((int)(ld_acquire_sys(a.my_flags + q) - e) < 0) __nanosleep(64);
    }
  }
  if (a.backward && threadIdx.x == 32) {
    while ((int)(ld_acquire_sys(a.ctrl + 2) - e) < 0) __nanosleep(32);  // g_own is completely initialised
  }
  __syncthreads();
  // ---- phase 3 (backward): add the returned ghost gradients into the owned rows
  if (a.backward) {
    float* d = reinterpret_cast<float*>(a.dst);
    for (long long i = gtid; i < a.n_send * C4; i += gstride) {
      const long long r = i / C4;
      const int c = (int)(i - r * C4);
      const float4 v = a.back[i];
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + (a.send_idx[r] * C4 + c) * 4), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w)
                   : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(a.ctrl + 3, 1u);
    if (t == gridDim.x - 1) {
      a.ctrl[3] = 0u;
      st_release_sys(a.ctrl, e);  // the epoch of this call site advances once per launch
    }
  }
}
}  // namespace bsms

using namespace bsms;

extern "C" int bsms_ipc_alloc(size_t bytes, void** ptr_out) {
  BSMS_CHECK_ARG(ptr_out && bytes > 0, "bsms_ipc_alloc: bad argument");
  BSMS_CUDA(cudaMalloc(ptr_out, bytes));
  BSMS_CUDA(cudaMemset(*ptr_out, 0, bytes));
  BSMS_CUDA(cudaDeviceSynchronize());
  return BSMS_OK;
}
extern "C" int bsms_ipc_free(void* ptr) {
  BSMS_CUDA(cudaFree(ptr));
  return BSMS_OK;
}
extern "C" int bsms_ipc_export(void* ptr, uint8_t* handle64) {
  BSMS_CHECK_ARG(ptr && handle64, "bsms_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  BSMS_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return BSMS_OK;
}
extern "C" int bsms_ipc_open(const uint8_t* handle64, void** ptr_out) {
  BSMS_CHECK_ARG(ptr_out && handle64, "bsms_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  BSMS_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return BSMS_OK;
}
extern "C" int bsms_ipc_close(void* ptr) {
  BSMS_CUDA(cudaIpcCloseMemHandle(ptr));
  return BSMS_OK;
}

extern "C" int bsms_halo_exchange(const bsms_halo_args* h, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BSMS_CHECK_ARG(h != nullptr, "bsms_halo_exchange: null argument");
  BSMS_CHECK_ARG(h->world >= 1 && h->world <= kMaxWorld && h->rank >= 0 && h->rank < h->world, "bsms_halo_exchange: world 1..%d", kMaxWorld);
  BSMS_CHECK_ARG(h->channels >= 4 && h->channels % 4 == 0, "bsms_halo_exchange: channels must be a multiple of 4");
  BSMS_CHECK_ARG(h->src && h->dst && h->ctrl && h->my_flags, "bsms_halo_exchange: null buffer");
  HaloArgs a;
  a.world = h->world;
  a.rank = h->rank;
  a.C4 = h->channels / 4;
  a.backward = h->backward;
  a.n_own = h->n_own;
  a.n_ghost = h->n_ghost;
  a.n_send = h->n_send;
  a.src = (const float4*)h->src;
  a.dst = (float4*)h->dst;
  a.send_idx = (const long long*)h->send_idx;
  a.back = (const float4*)h->back;
  a.my_flags = (unsigned*)h->my_flags;
  a.ctrl = (unsigned*)h->ctrl;
  for (int q = 0; q <= kMaxWorld; ++q) {
    a.send_off[q] = q <= h->world ? h->send_off[q] : h->send_off[h->world];
    a.recv_off[q] = q <= h->world ? h->recv_off[q] : h->recv_off[h->world];
  }
  for (int q = 0; q < kMaxWorld; ++q) {
    a.peer_dst[q] = q < h->world ? (float4*)h->peer_dst[q] : nullptr;
    a.peer_flag[q] = q < h->world ? (unsigned*)h->peer_flag[q] : nullptr;
  }
  BSMS_CHECK_ARG(a.send_off[h->world] == h->n_send && a.recv_off[h->world] == h->n_ghost, "bsms_halo_exchange: offsets do not sum up");
  int dev = 0, sms = 148;
  BSMS_CUDA(cudaGetDevice(&dev));
  BSMS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long work = (h->n_own + h->n_send + h->n_ghost) * a.C4;
  const int grid = (int)std::max<long long>(1, std::min<long long>((work + 255) / 256, 2ll * sms));
  ProfScope ps_(PK_TRANSFER, st);
  k_halo<<<grid, 256, 0, st>>>(a);
  BSMS_LAUNCHED();
  return BSMS_OK;
}
