"""Test infrastructure: compile the reference's two CUDA loss extensions (emd_linear/, chamfer_extension/) from the
sources where they lie under /root/reference into oracle/_ref/cuda_ext/ (git-ignored, travels to the GPU box), for
sm_100a.  Nothing of the reference is copied into the repository; the product never loads these files -- only
tests/test_losses_ref_gpu.py does, to compare the B200 kernels against the reference's own kernels on the same inputs.

    python oracle/build_ref_cuda.py          # needs /root/reference; a no-op where it is absent
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "cuda_ext")
REF = os.environ.get("CTB_REFERENCE_ROOT", "/root/reference")
EXTS = {"ref_emd": ("emd_linear", ["emd.cpp", "emd_cuda.cu"]),
        "ref_chamfer": ("chamfer_extension", ["chamfer_cuda.cpp", "chamfer.cu"])}


def built(name):
    return os.path.isfile(os.path.join(OUT, name, name + ".so"))


def build(verbose=False):
    if not os.path.isdir(REF):
        return {}
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils import cpp_extension
    done = {}
    for name, (sub, files) in EXTS.items():
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        if built(name):
            done[name] = "cached"
            continue
        try:
            cpp_extension.load(name=name, sources=[os.path.join(REF, sub, f) for f in files], build_directory=bdir,
                               extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"], verbose=verbose,
                               is_python_module=False)
            done[name] = "built"
        except Exception as exc:           # the recipe records why a reference file does not compile; never fatal
            done[name] = "failed: %s" % (str(exc).strip().splitlines()[-1][:300] if str(exc).strip() else repr(exc))
    return done


def load(name):
    """Import a built reference extension (GPU box: the prebuilt .so only)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols)
    path = os.path.join(OUT, name, name + ".so")
    if not os.path.isfile(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
