"""Host-side mirror of the reference's emd_linear/emd_module.py over the C ABI.

Same names and call: emdFunction.apply(xyz1, xyz2, eps, iters) -> (dist [B, n], assignment [B, n] int32), emdModule
(emd_module.py:30-87; train_inpainter.py:187-189 takes torch.sqrt(dist).mean(1).mean()).  Gradient for xyz1 only, like
the reference.  One kernel launch runs all iterations (csrc/ctb_emd.cuh); none of the reference's eleven scratch
tensors is allocated up to 4096 points (above, one workspace).  GPU tensors only; any n <= 32768 (the reference: multiples of 1024), any batch size."""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _call, _ptr, _require_cuda, _stream


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        _require_cuda(xyz1, xyz2)
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        assert n == m
        assert xyz1.size()[0] == xyz2.size()[0]
        assert xyz1.size(2) == 3 and xyz2.size(2) == 3
        assert int(iters) >= 1
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        dist = torch.empty(batchsize, n, device=xyz1.device)
        assignment = torch.empty(batchsize, n, device=xyz1.device, dtype=torch.int32)
        nbytes = _lib.load().ctb_emd_workspace_bytes(batchsize, n)          # 0 up to 4096 points
        ws = torch.empty(nbytes, dtype=torch.uint8, device=xyz1.device) if nbytes else None
        with torch.cuda.device(xyz1.device):
            _call("ctb_emd_fwd", _ptr(xyz1), _ptr(xyz2), _ptr(dist), _ptr(assignment), _ptr(ws), ctypes.c_size_t(nbytes),
                  batchsize, n, ctypes.c_float(eps), int(iters), _stream(xyz1))
        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    @once_differentiable
    def backward(ctx, graddist, gradidx):
        xyz1, xyz2, assignment = ctx.saved_tensors
        graddist = graddist.contiguous().float()
        gradxyz1 = torch.empty_like(xyz1)
        with torch.cuda.device(xyz1.device):
            _call("ctb_emd_bwd", _ptr(xyz1), _ptr(xyz2), _ptr(graddist), _ptr(assignment), _ptr(gradxyz1), xyz1.size(0),
                  xyz1.size(1), _stream(xyz1))
        return gradxyz1, torch.zeros_like(xyz2), None, None


class emdModule(nn.Module):
    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)
