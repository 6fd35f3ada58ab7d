"""SyncBatchNorm with the cross-GPU exchange fused into its kernels.

    model = ctb.syncbn.convert_sync_batchnorm(model)          # instead of torch.nn.SyncBatchNorm.convert_sync_batchnorm
    model = DistributedDataParallel(model, device_ids=[local_rank])

Same parameters / buffers / state_dict keys as nn.BatchNorm*d and nn.SyncBatchNorm (train_classification.py:107-109 wraps
the model exactly like that), so checkpoints are interchangeable.  In training mode every layer runs two kernels per
direction (csrc/ctb_syncbn.cuh): the statistics kernel stores its per-channel partial sums straight into the exchange
block of EVERY rank over NVLink peer memory and publishes an epoch flag; the elementwise kernel waits for all ranks'
flags and sums the partials in rank order.  No NCCL call, no host synchronisation; the kernels capture into a CUDA graph.
torch.distributed is plumbing here: its symmetric-memory rendezvous hands out the peer pointers once, at conversion.

Eval mode (running statistics) and world_size 1 fall through to F.batch_norm.  Equal per-rank batch shapes are assumed
(weak scaling: fixed per-GPU batch), as in the reference's DistributedSampler set-up.
"""
import ctypes

import torch
import torch.distributed as dist
from torch import nn
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _call, _ptr, _stream


class _ExchangePool:
    """One symmetric-memory allocation for all layers of a model: per layer and direction a block
    data f32 [2][W][2 C] + flag u32 [2][W], plus local epoch / done counters."""

    def __init__(self, blocks, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        W = self.world
        offs, total = [], 0
        for C in blocks:
            data_bytes = 2 * W * 2 * C * 4
            flag_bytes = 2 * W * 4
            offs.append((total, total + data_bytes))
            total += (data_bytes + flag_bytes + 255) // 256 * 256
        self.buf = symm_mem.empty(max(total, 256), dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        torch.cuda.synchronize(device)
        dist.barrier(group)                       # every rank has zeroed its block before anyone writes into it
        peers = [int(p) for p in self.handle.buffer_ptrs]
        # device arrays of peer pointers, one pair per block: [n_blocks][2][W] int64
        table = torch.empty((len(blocks), 2, W), dtype=torch.int64)
        for i, (d_off, f_off) in enumerate(offs):
            for r in range(W):
                table[i, 0, r] = peers[r] + d_off
                table[i, 1, r] = peers[r] + f_off
        self.table = table.to(device)
        self.local = torch.zeros((len(blocks), 4), dtype=torch.int32, device=device)     # epoch [1] | done [2] | pad
        lib = _lib.load()
        s_offs, s_total = [], 0                    # chunk partials + chunk counters of the statistics kernels
        for C in blocks:
            s_offs.append(s_total)
            s_total += (int(lib.ctb_syncbn_scratch_bytes(C)) + 15) // 16 * 16
        self.scratch = torch.zeros(max(s_total, 16), dtype=torch.uint8, device=device)
        self.structs = []
        for i in range(len(blocks)):
            st = _lib.CtbBnExchange(
                ctypes.c_void_p(self.table[i, 0].data_ptr()), ctypes.c_void_p(self.table[i, 1].data_ptr()),
                ctypes.c_void_p(self.local[i].data_ptr()), ctypes.c_void_p(self.local[i].data_ptr() + 4),
                ctypes.c_void_p(self.scratch.data_ptr() + s_offs[i]), self.rank, W)
            self.structs.append(st)


class _SyncBNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, ex_fwd, ex_bwd):
        xc = x.contiguous().float()
        B, C = xc.size(0), xc.size(1)
        L = xc.numel() // (B * C)
        y = torch.empty_like(xc)
        mean = torch.empty(C, dtype=torch.float32, device=xc.device)
        invstd = torch.empty(C, dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _call("ctb_syncbn_fwd", _ptr(xc), _ptr(weight), _ptr(bias), _ptr(y), _ptr(mean), _ptr(invstd), _ptr(running_mean),
                  _ptr(running_var), ctypes.byref(ex_fwd), B, C, L, ctypes.c_float(eps), ctypes.c_float(momentum), _stream(xc))
        ctx.save_for_backward(xc, weight, mean, invstd)
        ctx.ex_bwd = ex_bwd
        ctx.has_bias = bias is not None
        ctx.in_dtype = x.dtype
        # same shape as x; not a view: in-place ReLU may follow.  Arithmetic is fp32; other input dtypes are converted
        # on the way in and out, like nn.SyncBatchNorm returns its input's dtype.
        return y if x.dtype == torch.float32 else y.to(x.dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        xc, weight, mean, invstd = ctx.saved_tensors
        B, C = xc.size(0), xc.size(1)
        L = xc.numel() // (B * C)
        gyc = gy.contiguous().float()
        gx = torch.empty_like(xc)
        gw = torch.empty(C, dtype=torch.float32, device=xc.device) if weight is not None else None
        gb = torch.empty(C, dtype=torch.float32, device=xc.device) if ctx.has_bias else None
        with torch.cuda.device(xc.device):
            _call("ctb_syncbn_bwd", _ptr(xc), _ptr(gyc), _ptr(weight), _ptr(mean), _ptr(invstd), _ptr(gx), _ptr(gw), _ptr(gb),
                  ctypes.byref(ctx.ex_bwd), B, C, L, _stream(xc))
        if ctx.in_dtype != torch.float32:
            gx = gx.to(ctx.in_dtype)
        return gx, gw, gb, None, None, None, None, None, None


class CtbSyncBatchNorm(nn.modules.batchnorm._BatchNorm):
    """Drop-in for nn.SyncBatchNorm (any input rank >= 2, channels on dim 1)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        self._ex = None            # (forward exchange, backward exchange), set by convert_sync_batchnorm

    def _check_input_dim(self, input):
        if input.dim() < 2:
            raise ValueError("expected at least 2D input (got %dD input)" % input.dim())

    def forward(self, input):
        self._check_input_dim(input)
        use_batch_stats = self.training or not self.track_running_stats
        if not use_batch_stats or self._ex is None or not input.is_cuda:
            return nn.functional.batch_norm(input, self.running_mean, self.running_var, self.weight, self.bias,
                                            use_batch_stats, self.momentum if self.momentum is not None else 0.0, self.eps)
        momentum = self.momentum
        if self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if momentum is None:
                momentum = 1.0 / float(self.num_batches_tracked)
        rm = self.running_mean if self.track_running_stats else None
        rv = self.running_var if self.track_running_stats else None
        return _SyncBNFn.apply(input, self.weight, self.bias, rm, rv, self.eps, momentum if momentum is not None else 0.0,
                               self._ex[0], self._ex[1])


def convert_sync_batchnorm(module, process_group=None, device=None):
    """Replace every nn.BatchNorm*d / nn.SyncBatchNorm of `module` (already on its GPU) by CtbSyncBatchNorm sharing the
    original parameters and buffers, and allocate the peer-mapped exchange blocks (collective call: every rank of
    `process_group` must convert the same model).  world_size 1 or no process group: returns the module unchanged."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return module
    layers = []

    def swap(parent):
        for name, child in list(parent.named_children()):
            if isinstance(child, nn.modules.batchnorm._BatchNorm) and not isinstance(child, CtbSyncBatchNorm):
                new = CtbSyncBatchNorm(child.num_features, child.eps, child.momentum, child.affine, child.track_running_stats)
                if child.affine:
                    new.weight, new.bias = child.weight, child.bias
                if child.track_running_stats:
                    new.running_mean, new.running_var = child.running_mean, child.running_var
                    new.num_batches_tracked = child.num_batches_tracked
                new.training = child.training
                setattr(parent, name, new)
                layers.append(new)
            else:
                swap(child)

    if isinstance(module, nn.modules.batchnorm._BatchNorm):
        raise ValueError("convert_sync_batchnorm expects a container module")
    swap(module)
    if not layers:
        return module
    if device is None:
        device = next(module.parameters()).device
    blocks = []
    for m in layers:
        blocks += [m.num_features, m.num_features]         # forward and backward exchange
    pool = _ExchangePool(blocks, device, process_group)
    for i, m in enumerate(layers):
        m._ex = (pool.structs[2 * i], pool.structs[2 * i + 1])
        m._pool = pool                                     # keeps the symmetric buffer alive
    return module
