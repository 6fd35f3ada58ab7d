"""Host-side mirror of the reference's chamfer_extension/dist_chamfer.py over the C ABI.

Same names and semantics: ChamferFunction (forward -> dist1 [B,n], dist2 [B,m]; backward via the saved nearest-neighbour
indices), ChamferDist, loss_chamfer, loss_chamfer_adj, loss_chamder_2d (dist_chamfer.py:10-98).  GPU tensors only, like
the reference; no CPU fallback."""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _call, _ptr, _require_cuda, _stream


class ChamferFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        _require_cuda(xyz1, xyz2)
        assert xyz1.device == xyz2.device
        assert xyz1.size(0) == xyz2.size(0)
        assert xyz1.size(2) == 3 and xyz2.size(2) == 3
        a, b = xyz1.contiguous().float(), xyz2.contiguous().float()
        B, n, _ = a.shape
        m = b.size(1)
        dev = a.device
        dist1 = torch.empty(B, n, device=dev)
        dist2 = torch.empty(B, m, device=dev)
        idx1 = torch.empty(B, n, dtype=torch.int32, device=dev)
        idx2 = torch.empty(B, m, dtype=torch.int32, device=dev)
        nbytes = _lib.load().ctb_chamfer_workspace_bytes(B, n, m)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _call("ctb_chamfer_fwd", _ptr(a), _ptr(b), _ptr(dist1), _ptr(dist2), _ptr(idx1), _ptr(idx2), _ptr(ws),
                  ctypes.c_size_t(nbytes), B, n, m, _stream(a))
        ctx.save_for_backward(a, b, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    @once_differentiable
    def backward(ctx, graddist1, graddist2, _gi1, _gi2):
        a, b, idx1, idx2 = ctx.saved_tensors
        B, n, _ = a.shape
        m = b.size(1)
        g1, g2 = torch.empty_like(a), torch.empty_like(b)
        gd1, gd2 = graddist1.contiguous().float(), graddist2.contiguous().float()
        with torch.cuda.device(a.device):
            _call("ctb_chamfer_bwd", _ptr(a), _ptr(b), _ptr(gd1), _ptr(gd2), _ptr(idx1), _ptr(idx2), _ptr(g1), _ptr(g2),
                  B, n, m, _stream(a))
        return g1, g2


class ChamferDist(nn.Module):
    def forward(self, input1, input2):
        dist1, dist2, _, _ = ChamferFunction.apply(input1, input2)
        return dist1, dist2


def loss_chamfer(pc_1, pc_2):
    """pc [B, 3, 1, N] (dist_chamfer.py:67-77)"""
    dist_1, dist_2 = ChamferDist()(pc_1[:, :, 0].permute(0, 2, 1).contiguous(), pc_2[:, :, 0].permute(0, 2, 1).contiguous())
    return torch.mean(dist_1) + torch.mean(dist_2)


def loss_chamfer_adj(pc_1, pc_2):
    """the PCN variant (dist_chamfer.py:81-90)"""
    dist_1, dist_2 = ChamferDist()(pc_1[:, :, 0].permute(0, 2, 1).contiguous(), pc_2[:, :, 0].permute(0, 2, 1).contiguous())
    return (torch.mean(torch.sqrt(dist_1)) + torch.mean(torch.sqrt(dist_2))) / 2


def loss_chamder_2d(pc_1, pc_2):
    """2-D clouds padded with a zero coordinate (dist_chamfer.py:93-98; the reference's spelling is kept)"""
    zeros_1 = torch.zeros(pc_1.size(0), 1, 1, pc_1.size(-1), device=pc_1.device)
    zeros_2 = torch.zeros(pc_2.size(0), 1, 1, pc_2.size(-1), device=pc_1.device)
    return loss_chamfer(torch.cat([pc_1, zeros_1], dim=1), torch.cat([pc_2, zeros_2], dim=1))
