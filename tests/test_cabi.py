"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/bsms_b200.h declares; no compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bsms_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bsms_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ["bsms_plan_build", "bsms_cal_ew", "bsms_edge_conv", "bsms_conv_down_pool", "bsms_unpool_conv_up",
                 "bsms_unpool_rows", "bsms_gmp_forward", "bsms_gmp_backward"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "bsms_gnn_b200", "libbsms_b200.so")
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/bsms_b200.h but not exported"
    lib.bsms_version.restype = ctypes.c_int
    assert lib.bsms_version() >= 100
    lib.bsms_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.bsms_last_error(), bytes)


def test_python_binding_covers_header():
    from bsms_gnn_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_no_cpu_fallback():
    import torch
    from bsms_gnn_b200 import _lib
    from bsms_gnn_b200.ops import BSGMP, GMP, WeightedEdgeConv
    m = BSGMP(1, 128, 3, 2)
    g = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(_lib.BsmsError):
        m(torch.zeros(2, 128), [torch.tensor([0])], [g, torch.zeros(2, 0, dtype=torch.long)], torch.zeros(2, 2))
    with pytest.raises(_lib.BsmsError):
        GMP(128, 3, 2)(torch.zeros(2, 128), g, torch.zeros(2, 2))
    with pytest.raises(_lib.BsmsError):
        WeightedEdgeConv()(torch.zeros(2, 128), g, torch.ones(2))
    with pytest.raises(_lib.BsmsError):
        GMP(64, 3, 2)  # only the latent width of the reference configs is built


def test_state_dict_keys_match_reference_layout():
    from bsms_gnn_b200.ops import BSGMP
    from oracle import bsms_oracle as O
    m = BSGMP(2, 128, 3, 3)
    ref_keys = set(O.init_params(2, pos_dim=3).keys())
    assert set(m.state_dict().keys()) == ref_keys
    sd = m.state_dict()
    assert tuple(sd["down_gmps.0.mlp_edge.seq.0.weight"].shape) == (128, 260)
    assert tuple(sd["bottom_gmp.mlp_node.seq.0.weight"].shape) == (128, 256)
