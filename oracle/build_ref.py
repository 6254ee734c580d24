"""Vendor the UNMODIFIED reference hot path into oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BASELINE INFRASTRUCTURE — never imported by the product package.

`/root/reference` exists only in the build container; the GPU box gets whatever lies under the repo
snapshot.  This recipe copies the reference's own Python modules for the path (the processor
`src/ops`, its caller `src/models`, the scatter helpers and normaliser in `src/utils`, the
hierarchy builder `src/graph_wrappers`, and `src/trainer` for the step semantics) byte for byte into
`oracle/_ref/src/`, plus a manifest with their sha256, so that

  * `bench.py --impl reference` and bench.py's `cpu_baseline` time the reference's OWN modules on
    the box's host cores (`cpu_baseline.kind = "reference"`), and
  * the `-m gpu` drop-in test drives the unmodified `BSMS_Simulator.forward` with
    `bsms_gnn_b200.ops` swapped in behind it.

No file of the reference is committed: `oracle/_ref/` is listed in .gitignore.  Run by
`__graft_entry__.build()` whenever /root/reference is present.
"""
import hashlib
import json
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
PACKAGES = ["ops", "models", "utils", "graph_wrappers", "trainer"]


def build(verbose: bool = False) -> bool:
    src_root = os.path.join(REF, "src")
    if not os.path.isdir(src_root):
        return False
    manifest = {}
    for pkg in PACKAGES:
        sdir, ddir = os.path.join(src_root, pkg), os.path.join(DST, "src", pkg)
        os.makedirs(ddir, exist_ok=True)
        for name in sorted(os.listdir(sdir)):
            if not name.endswith(".py"):
                continue
            shutil.copyfile(os.path.join(sdir, name), os.path.join(ddir, name))
            with open(os.path.join(ddir, name), "rb") as f:
                manifest[f"src/{pkg}/{name}"] = hashlib.sha256(f.read()).hexdigest()
    commit = None
    sub = os.path.join(REF, ".SUBMODULES.json")
    if os.path.exists(sub):
        try:
            commit = json.load(open(sub))
        except Exception:
            commit = None
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "submodules": commit, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files vendored")
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    if not ok:
        print("reference tree not present; nothing vendored")
