"""Import the UNMODIFIED reference hot path (CPU or, with the product ops swapped in, the caller side).

TEST / BASELINE INFRASTRUCTURE ONLY.  Source of the modules, in this order:
  1. /root/reference/src            — the build container;
  2. oracle/_ref/src                — the byte-for-byte copy made by oracle/build_ref.py (git-ignored,
                                      travels to the GPU box with the repo snapshot).
Used by tests/golden/make_golden.py (golden vectors), by the tests that pin oracle/bsms_oracle.py
against the real reference, by bench.py's reference arm / cpu_baseline, and by the drop-in test
that runs the reference's own `BSMS_Simulator.forward` on top of `bsms_gnn_b200.ops`.
Recipe: SURVEY.md Appendix B (stubs for modules that are off the arithmetic path).
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = ["/root/reference/src", os.path.join(HERE, "_ref", "src")]
REF_PACKAGES = ("utils", "ops", "models", "graph_wrappers", "trainer")


def ref_src():
    for c in CANDIDATES:
        if os.path.isdir(os.path.join(c, "ops")):
            return c
    return None


def available() -> bool:
    return ref_src() is not None


def _stubs():
    for name in ["matplotlib", "matplotlib.pyplot", "pytz", "torchsummary"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].figure = types.SimpleNamespace(Figure=object)
    sys.modules["pytz"].timezone = lambda s: None
    sys.modules["torchsummary"].summary = lambda *a, **k: None
    if "sparse_dot_mkl" not in sys.modules:
        sdm = types.ModuleType("sparse_dot_mkl")
        # only the sparsity pattern of (A+I)^2 is consumed (bsms_graph_wrapper.py:100-102)
        sdm.dot_product_mkl = lambda a, b: (a @ b).tocsr()
        sys.modules["sparse_dot_mkl"] = sdm
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:  # logging only (utils/basic.py:14)
            sys.modules["wandb"] = types.ModuleType("wandb")
    if "tabulate" not in sys.modules:
        try:
            import tabulate  # noqa: F401
        except Exception:
            t = types.ModuleType("tabulate")
            t.tabulate = lambda *a, **k: ""
            sys.modules["tabulate"] = t


def _purge():
    for name in list(sys.modules):
        if name.split(".")[0] in REF_PACKAGES:
            mod = sys.modules[name]
            f = getattr(mod, "__file__", "") or ""
            if any(f.startswith(c) for c in CANDIDATES):
                del sys.modules[name]


def load():
    """Returns the reference modules (ops, graph_wrappers, utils, ...) or raises ImportError."""
    src = ref_src()
    if src is None:
        raise ImportError("reference tree not present (neither /root/reference nor oracle/_ref)")
    if src not in sys.path:
        sys.path.insert(0, src)
    _stubs()
    import utils  # noqa: F401
    import ops
    import graph_wrappers
    import utils.mesh_convertions as mesh_convertions

    return types.SimpleNamespace(ops=ops, graph_wrappers=graph_wrappers, utils=utils,
                                 mesh_convertions=mesh_convertions, src=src,
                                 kind="reference" if src.startswith("/root/reference") else "reference (oracle/_ref copy)")


def load_simulator(swap_ops=None, device=None):
    """The reference's `models.model` module, unmodified.

    swap_ops: a module exporting `MLP` and `BSGMP` (e.g. bsms_gnn_b200.ops) that takes the place of the
    reference's `ops` package for `from ops import MLP, BSGMP` (src/models/model.py:2) — the one-line swap
    INTEGRATION.md documents.  None keeps the reference's own ops.
    device: overrides the module-level `device` globals of models.model / utils.normalizer (they pick cuda
    whenever it is available, model.py:5).
    """
    src = ref_src()
    if src is None:
        raise ImportError("reference tree not present (neither /root/reference nor oracle/_ref)")
    if src not in sys.path:
        sys.path.insert(0, src)
    _stubs()
    import utils  # noqa: F401
    saved_ops = sys.modules.get("ops")
    sys.modules.pop("models", None)
    sys.modules.pop("models.model", None)
    try:
        if swap_ops is not None:
            shim = types.ModuleType("ops")
            shim.MLP, shim.BSGMP = swap_ops.MLP, swap_ops.BSGMP
            sys.modules["ops"] = shim
        else:
            import ops  # noqa: F401
        model_mod = importlib.import_module("models.model")
    finally:
        if saved_ops is not None:
            sys.modules["ops"] = saved_ops
        else:
            sys.modules.pop("ops", None)
        # the next load_simulator call must re-import models.model against ITS ops
        sys.modules.pop("models", None)
        sys.modules.pop("models.model", None)
    if device is not None:
        import torch
        model_mod.device = torch.device(device)
    return model_mod
