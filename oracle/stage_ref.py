"""Stage the UNMODIFIED reference files the parity tests and the reference bench arm execute into oracle/_ref/.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is Python, so "building" it means copying the files of the path
(and of its callers) verbatim from where they lie under /root/reference:
    layers/  model_zoo/  unet2d/  utils/  configs/
`oracle/_ref/` is git-ignored (no reference source enters the history) but travels to the GPU box with the snapshot,
where /root/reference does not exist.  Run by `__graft_entry__.build()` whenever /root/reference is present:
    python oracle/stage_ref.py [reference_root]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SUBTREES = ("layers", "model_zoo", "unet2d", "utils", "configs")


def stage(ref_root="/root/reference", dst=DST):
    if not os.path.isfile(os.path.join(ref_root, "layers", "cloud_transform.py")):
        return False
    os.makedirs(dst, exist_ok=True)
    for sub in SUBTREES:
        src = os.path.join(ref_root, sub)
        if not os.path.isdir(src):
            continue
        out = os.path.join(dst, sub)
        if os.path.isdir(out):
            shutil.rmtree(out)
        shutil.copytree(src, out, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(dst, "STAGED_FROM"), "w") as f:
        f.write(ref_root + "\n")
    return True


if __name__ == "__main__":
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("staged" if ok else "reference not found", DST)
