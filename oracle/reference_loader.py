"""Import the UNMODIFIED reference (`/root/reference/layers/*`) under two dependency shims -- TEST
INFRASTRUCTURE ONLY, build-container only (the GPU box has no /root/reference; nothing under
tests -m gpu, smoke() or bench.py calls this at run time).

The reference cannot be imported as shipped here: `torch_scatter` (layers/cloud_transform.py:7) and
`pytorch3d` (layers/utils.py:6) are not installed and there is no network.  Both are third-party
dependencies whose sources are not under /root/reference (install_deps.sh:6,10; versions unpinned /
py36_cu101_pyt160 wheel).  The shims restate their published semantics:

* torch_scatter.scatter_max(src, index, dim, out) -> (out, arg): `out` is NOT re-initialised (the
  zeros act as a floor), index is broadcast to src, update rule in ascending e is
  `if src > out: out = src; arg = e`, arg sentinel = src.size(dim); backward scatters grad_out to the
  single arg winner (ScatterMax::backward: zeros[..., E+1].scatter_(dim, arg, grad).narrow(0, E)).
  NOTE scatter_reduce('amax') has a different backward (splits ties), hence the explicit Function.
* pytorch3d.transforms.so3.so3_exponential_map(log_rot[H,3]) -> [H,3,3]: Rodrigues formula with the
  squared norm clamped to >= 1e-4.

Used by tests/golden/make_golden.py to mint the golden fixtures and by tests (skipped when
/root/reference is absent) to validate the oracle restatements against the real reference code.
"""
import importlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref")          # verbatim copy made by oracle/stage_ref.py (travels to the GPU box)
REPO_ROOT = os.path.dirname(_HERE)
DROPIN_ROOT = os.path.join(REPO_ROOT, "dropin")


def _pick_root():
    env = os.environ.get("CTB_REFERENCE_ROOT")
    for cand in (env, "/root/reference", STAGED_ROOT):
        if cand and os.path.isfile(os.path.join(cand, "layers", "cloud_transform.py")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _pick_root()


class _ScatterMaxFirst(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, out):
        E = src.size(-1)
        C = out.size(-1)
        index_b = index.expand_as(src)
        out.scatter_reduce_(-1, index_b, src, reduce="amax", include_self=True)
        at = out.gather(-1, index_b)
        cand = (src == at) & (src > 0)
        e = torch.arange(E, dtype=torch.int64, device=src.device).expand_as(src)
        arg = torch.full(out.shape, E, dtype=torch.int64, device=src.device)
        arg.scatter_reduce_(-1, index_b, torch.where(cand, e, torch.full_like(e, E)), reduce="amin",
                            include_self=True)
        ctx.save_for_backward(arg)
        ctx.E = E
        ctx.mark_non_differentiable(arg)
        ctx.mark_dirty(out)
        return out, arg

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        E = ctx.E
        g = torch.zeros(*grad_out.shape[:-1], E + 1, dtype=grad_out.dtype, device=grad_out.device)
        g.scatter_(-1, arg, grad_out)
        return g[..., :E], None, None


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    assert out is not None, "the reference always passes out= (cloud_transform.py:171)"
    assert dim in (-1, src.dim() - 1)
    return _ScatterMaxFirst.apply(src, index, out)


def so3_exponential_map(log_rot, eps: float = 0.0001):
    nrms = (log_rot * log_rot).sum(1)
    theta = torch.clamp(nrms, eps).sqrt()
    fac1 = theta.sin() / theta
    fac2 = (1.0 - theta.cos()) / (theta * theta)
    K = torch.zeros(log_rot.shape[0], 3, 3, dtype=log_rot.dtype, device=log_rot.device)
    x, y, z = log_rot.unbind(1)
    K[:, 0, 1], K[:, 0, 2] = -z, y
    K[:, 1, 0], K[:, 1, 2] = z, -x
    K[:, 2, 0], K[:, 2, 1] = -y, x
    return fac1[:, None, None] * K + fac2[:, None, None] * torch.bmm(K, K) + \
        torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)[None]


def install_shims():
    if "torch_scatter" not in sys.modules:
        m = types.ModuleType("torch_scatter")
        m.scatter_max = scatter_max
        sys.modules["torch_scatter"] = m
    if "pytorch3d" not in sys.modules:
        p = types.ModuleType("pytorch3d")
        t = types.ModuleType("pytorch3d.transforms")
        s = types.ModuleType("pytorch3d.transforms.so3")
        s.so3_exponential_map = so3_exponential_map
        t.so3 = s
        p.transforms = t
        sys.modules["pytorch3d"] = p
        sys.modules["pytorch3d.transforms"] = t
        sys.modules["pytorch3d.transforms.so3"] = s


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "layers", "cloud_transform.py"))


def load_reference_layers():
    """returns the reference's `layers.cloud_transform`, `layers.utils`, `layers.multihead_ct` modules."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    install_shims()
    # the reference package is named `layers`; make sure no other `layers` shadows it
    for k in [k for k in sys.modules if k == "layers" or k.startswith("layers.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ct = importlib.import_module("layers.cloud_transform")
        ut = importlib.import_module("layers.utils")
        mh = importlib.import_module("layers.multihead_ct")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    assert os.path.realpath(ct.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    return ct, ut, mh


def _purge(prefixes=("layers", "unet2d", "utils", "model_zoo")):
    for k in [k for k in sys.modules if any(k == p or k.startswith(p + ".") for p in prefixes)]:
        del sys.modules[k]


class reference_tree:
    """Context manager: the reference's package tree (`layers`, `unet2d`, `utils`, ...) importable, with the two
    dependency shims installed.  dropin=True puts this repo's `dropin/` FIRST on sys.path: the reference's `layers` is
    a namespace package (no __init__.py), so `layers.cloud_transform` then resolves to dropin/layers/cloud_transform.py
    (the B200 kernels) while every other module -- layers.multihead_ct*, layers.utils, unet2d.*, the model files --
    is the reference's own, unmodified file.  That is exactly what a user does to switch."""

    def __init__(self, dropin=False, root=None):
        self.dropin = dropin
        self.root = root or REFERENCE_ROOT

    def __enter__(self):
        if not os.path.isfile(os.path.join(self.root, "layers", "cloud_transform.py")):
            raise RuntimeError("reference tree not present at %s (run oracle/stage_ref.py in the build container)" % self.root)
        install_shims()
        _purge()
        self.added = ([DROPIN_ROOT] if self.dropin else []) + [self.root]
        if self.dropin and REPO_ROOT not in sys.path:
            sys.path.insert(0, REPO_ROOT)
        for p in reversed(self.added):
            sys.path.insert(0, p)
        return self

    def __exit__(self, *exc):
        for p in self.added:
            if p in sys.path:
                sys.path.remove(p)
        _purge()
        return False

    def load_model(self, rel_path, **params):
        """exec the model file like utils/train_util.py:23-27 get_model() and instantiate its `Model`."""
        env = {}
        with open(os.path.join(self.root, rel_path), "r") as f:
            exec(compile(f.read(), os.path.join(self.root, rel_path), "exec"), env)
        return env["Model"](**params)
