"""ctypes binding of libbsms_b200.so (the C-ABI declared in include/bsms_b200.h).

There is no CPU fallback: importing this module without the built library raises, and every call
site checks for CUDA tensors.  Build with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C bsms_gnn_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbsms_b200.so")

MODE_FP32, MODE_FP16X3, MODE_BF16 = 0, 1, 2
MODES = {"fp32": MODE_FP32, "fp16x3": MODE_FP16X3, "bf16": MODE_BF16}


class BsmsError(RuntimeError):
    pass


class LevelPlanC(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_edges", C.c_int32),
                ("src_d", C.c_void_p), ("dst_d", C.c_void_p), ("rowptr_d", C.c_void_p), ("perm_d", C.c_void_p),
                ("src_s", C.c_void_p), ("dst_s", C.c_void_p), ("rowptr_s", C.c_void_p), ("s2d", C.c_void_p)]


class GmpWeightsC(C.Structure):
    _fields_ = [("w_edge", C.c_void_p * 4), ("b_edge", C.c_void_p * 4),
                ("w_node", C.c_void_p * 4), ("b_node", C.c_void_p * 4)]


class HaloArgsC(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("channels", C.c_int32), ("backward", C.c_int32),
                ("n_own", C.c_int64), ("n_ghost", C.c_int64), ("n_send", C.c_int64),
                ("src", C.c_void_p), ("dst", C.c_void_p), ("send_idx", C.c_void_p),
                ("send_off", C.c_int32 * 9), ("recv_off", C.c_int32 * 9), ("peer_dst", C.c_void_p * 8),
                ("back", C.c_void_p), ("my_flags", C.c_void_p), ("peer_flag", C.c_void_p * 8), ("ctrl", C.c_void_p)]


EXPORTS = [
    "bsms_last_error", "bsms_version", "bsms_device_info", "bsms_plan_workspace_bytes", "bsms_plan_build", "bsms_fingerprint",
    "bsms_cal_ew", "bsms_permute_ew", "bsms_edge_conv", "bsms_conv_down_pool", "bsms_unpool_conv_up",
    "bsms_gather_rows", "bsms_unpool_rows", "bsms_gmp_workspace_bytes", "bsms_gmp_saved_bytes", "bsms_gmp_forward",
    "bsms_gmp_backward", "bsms_launch_count", "bsms_prof_enable", "bsms_prof_collect",
    "bsms_debug_edge_stage", "bsms_debug_lin_split", "bsms_masked_rmse", "bsms_clip_adamw_step", "bsms_inject_noise",
    "bsms_gmp_packed_bytes", "bsms_gmp_pack", "bsms_gmp_forward_packed",
    "bsms_components_host", "bsms_bistride_level_host", "bsms_host_free",
    "bsms_hierarchy_build_host", "bsms_hierarchy_level_host", "bsms_hierarchy_free_host",
    "bsms_ipc_alloc", "bsms_ipc_free", "bsms_ipc_export", "bsms_ipc_open", "bsms_ipc_close", "bsms_halo_exchange",
    "bsms_encode_in", "bsms_dense128_packed_bytes", "bsms_dense128_pack", "bsms_dense128_stack", "bsms_decode_out",
    "bsms_set_deterministic", "bsms_get_deterministic",
]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library is the product and there is no CPU fallback. "
            "Build it with __graft_entry__.build() (nvcc, sm_100a).")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    P = C.POINTER
    lib.bsms_last_error.restype = C.c_char_p
    lib.bsms_last_error.argtypes = []
    lib.bsms_version.restype = C.c_int
    lib.bsms_device_info.argtypes = [P(i64)]
    lib.bsms_plan_workspace_bytes.restype = sz
    lib.bsms_plan_workspace_bytes.argtypes = [i64, i64]
    lib.bsms_plan_build.argtypes = [vp, i64, i64, P(LevelPlanC), vp, vp, sz, vp]
    lib.bsms_fingerprint.argtypes = [P(vp), P(i64), i32, vp, vp]
    lib.bsms_cal_ew.argtypes = [P(LevelPlanC), vp, vp, vp, vp, vp, vp]
    lib.bsms_permute_ew.argtypes = [P(LevelPlanC), vp, vp, vp, vp]
    lib.bsms_edge_conv.argtypes = [P(LevelPlanC), vp, vp, vp, i32, i32, i32, vp]
    lib.bsms_conv_down_pool.argtypes = [P(LevelPlanC), vp, vp, i32, vp, vp, i32, i32, vp]
    lib.bsms_unpool_conv_up.argtypes = [P(LevelPlanC), vp, vp, i32, vp, vp, i32, i32, vp]
    lib.bsms_gather_rows.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp]
    lib.bsms_unpool_rows.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp]
    lib.bsms_gmp_workspace_bytes.restype = sz
    lib.bsms_gmp_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    lib.bsms_gmp_saved_bytes.restype = sz
    lib.bsms_gmp_saved_bytes.argtypes = [i32, i32]
    lib.bsms_gmp_forward.argtypes = [P(LevelPlanC), P(GmpWeightsC), vp, vp, i32, vp, vp, vp, i32, i32, i32, vp, sz, vp]
    lib.bsms_gmp_backward.argtypes = [P(LevelPlanC), P(GmpWeightsC), vp, vp, i32, vp, vp, vp, P(GmpWeightsC),
                                      i32, i32, i32, vp, sz, vp]
    lib.bsms_launch_count.restype = i64
    lib.bsms_launch_count.argtypes = []
    lib.bsms_set_deterministic.restype = None
    lib.bsms_set_deterministic.argtypes = [i32]
    lib.bsms_get_deterministic.restype = i32
    lib.bsms_get_deterministic.argtypes = []
    lib.bsms_debug_edge_stage.argtypes = [P(LevelPlanC), P(GmpWeightsC), vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    f64 = C.c_double
    lib.bsms_masked_rmse.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp]
    lib.bsms_clip_adamw_step.argtypes = [vp, vp, vp, vp, i64, vp, vp, f64, f64, f64, f64, f64, f64, f64, f64, i32, vp]
    lib.bsms_debug_lin_split.argtypes = [vp, i64, vp, i32, vp, i32, vp, vp, vp]
    lib.bsms_inject_noise.argtypes = [vp, i32, vp, i32, vp, i64, P(C.c_float), C.c_float, C.c_uint64, C.c_uint64, vp]
    lib.bsms_gmp_packed_bytes.restype = sz
    lib.bsms_gmp_packed_bytes.argtypes = []
    lib.bsms_gmp_pack.argtypes = [P(GmpWeightsC), i32, i32, vp, vp]
    lib.bsms_gmp_forward_packed.argtypes = [P(LevelPlanC), P(GmpWeightsC), vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, vp, sz, vp]
    lib.bsms_encode_in.argtypes = [vp, i64, i32, i32, i32, P(f64), P(f64), vp, vp, vp, vp, vp]
    lib.bsms_dense128_packed_bytes.restype = sz
    lib.bsms_dense128_packed_bytes.argtypes = [i32]
    lib.bsms_dense128_pack.argtypes = [P(vp), i32, i32, vp, vp]
    lib.bsms_dense128_stack.argtypes = [vp, i64, P(vp), P(vp), i32, i32, i32, i32, vp, vp, vp, vp]
    lib.bsms_decode_out.argtypes = [vp, i64, i32, i32, vp, vp, P(f64), P(f64), vp, vp, vp, vp, vp, i32, vp]
    lib.bsms_components_host.argtypes = [vp, i64, i64, vp, P(i64)]
    lib.bsms_bistride_level_host.argtypes = [vp, i64, i64, vp, i64, vp, vp, P(i64), P(vp), P(i64)]
    lib.bsms_host_free.argtypes = [vp]
    lib.bsms_host_free.restype = None
    lib.bsms_hierarchy_build_host.argtypes = [vp, i64, i64, vp, i32, i32, i32, P(vp)]
    lib.bsms_hierarchy_level_host.argtypes = [vp, i32, P(i64), P(i64), P(vp), P(vp)]
    lib.bsms_hierarchy_free_host.argtypes = [vp]
    lib.bsms_hierarchy_free_host.restype = None
    lib.bsms_ipc_alloc.argtypes = [sz, P(vp)]
    lib.bsms_ipc_free.argtypes = [vp]
    lib.bsms_ipc_export.argtypes = [vp, C.c_char_p]
    lib.bsms_ipc_open.argtypes = [C.c_char_p, P(vp)]
    lib.bsms_ipc_close.argtypes = [vp]
    lib.bsms_halo_exchange.argtypes = [P(HaloArgsC), vp]
    lib.bsms_prof_enable.argtypes = [C.c_int]
    lib.bsms_prof_collect.argtypes = [P(C.c_double), P(i64), C.c_int]
    for name in EXPORTS:
        getattr(lib, name)  # every symbol the header declares must resolve
    return lib


lib = _load()


def check(rc: int):
    if rc != 0:
        msg = lib.bsms_last_error().decode("utf-8", "replace")
        if rc == -3:
            raise IndexError(msg)
        raise BsmsError(f"libbsms_b200 error {rc}: {msg}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    """cudaStream_t of torch's current stream on the CURRENT device (call inside `torch.cuda.device(dev)`)."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise BsmsError("bsms_gnn_b200 runs on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                            f"got a tensor on {t.device}")


def launch_count() -> int:
    return int(lib.bsms_launch_count())


# bumped by anything that rewrites parameter storage behind autograd's back (train.FlatAdamW's raw kernel):
# part of the key of every cache of packed weight images
WEIGHTS_EPOCH = [0]

_WS = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream) — stream-ordered reuse."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


PROF_KINDS = ["edge_fwd_gemm", "node_fwd_gemm", "edge_combine", "ln_segsum", "dgrad", "wgrad", "ln_bwd",
              "edge_grad_segsum", "transfer", "other", "edge_chain", "edge_chain_bwd"]


def prof_enable(on: bool):
    check(lib.bsms_prof_enable(1 if on else 0))


def prof_collect():
    """-> {kind: (total_ms, launches)} since the last collect (synchronises the device)."""
    n = len(PROF_KINDS)
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    check(lib.bsms_prof_collect(ms, cnt, n))
    return {k: (ms[i], int(cnt[i])) for i, k in enumerate(PROF_KINDS)}


def device_info():
    out = (C.c_int64 * 4)()
    check(lib.bsms_device_info(out))
    return {"sms": out[0], "l2_bytes": out[1], "smem_optin": out[2], "cc": out[3]}
