"""bsms_gnn_b200 — B200-native (sm_100a) BSMS-GNN processor: drop-in for the reference's src/ops.

    from bsms_gnn_b200.ops import MLP, BSGMP, GMP, WeightedEdgeConv, Unpool

Importing the package loads libbsms_b200.so (C-ABI in include/bsms_b200.h); there is no CPU path.
`hierarchy` and `meshgen` are pure numpy/scipy preprocessing helpers and import without the library.
"""
__all__ = ["ops", "plan", "hierarchy", "meshgen"]
