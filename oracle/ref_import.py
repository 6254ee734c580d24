"""Import the UNMODIFIED reference hot path from /root/reference/src (CPU).

TEST INFRASTRUCTURE ONLY.  This works only in the build container (the GPU box has no
/root/reference); it is used by tests/golden/make_golden.py to generate the committed golden
vectors and by tests that pin oracle/bsms_oracle.py against the real reference when it is
present.  Recipe: SURVEY.md Appendix B (stubs for modules that are off the arithmetic path).
"""
import os
import sys
import types

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(REF_SRC)


def load():
    """Returns the reference modules (ops, graph_wrappers, utils) or raises ImportError."""
    if not available():
        raise ImportError("reference tree not present (expected on the GPU box)")
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
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
    import utils  # noqa: F401
    import ops
    import graph_wrappers
    import utils.mesh_convertions as mesh_convertions

    return types.SimpleNamespace(ops=ops, graph_wrappers=graph_wrappers, utils=utils,
                                 mesh_convertions=mesh_convertions)
