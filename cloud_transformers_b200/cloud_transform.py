"""Host-side mirror of the reference's `layers/cloud_transform.py` for the Splat / Slice hot path.

Same class names, constructor `(tensor_size=20, heads=4, dim=3)`, forward signatures, attributes and the
persistent `tensor_mod` buffer (so released checkpoints load strictly) as
  DifferentiableGridModule  layers/cloud_transform.py:29-59
  DifferentiablePositions   layers/cloud_transform.py:62-121
  Splat                     layers/cloud_transform.py:124-180
  Slice                     layers/cloud_transform.py:183-227
Underneath, each forward is one autograd.Function over the sm_100a C-ABI library (functional.py).

Fast path: when Splat / Slice receive the very tensors our DifferentiablePositions returned (as the
MHCT blocks do, layers/multihead_ct.py:99-107), they skip reading local_coordinate / flattened_index
from memory and recompute the positions from the keys inside the kernels (identical device code, so
identical bits); the gradient then flows straight to the keys.
"""
import torch
from torch import nn

from . import _lib
from . import functional as CF


class GradientBalancing(torch.autograd.Function):
    """layers/cloud_transform.py:12-23 -- scales the forward, leaves the gradient untouched.  Kept for API
    parity; the kernels fold this rule in (no (W-1)/2 factor in grad_keys)."""

    @staticmethod
    def forward(ctx, input, scale):
        return input * scale

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None


balance_op = GradientBalancing.apply


class DifferentiableGridModule(nn.Module):
    def __init__(self, tensor_size=20, heads=4, dim=3):
        '''
        :param tensor_size: spatial resolution of the feature map, tuple with len() == dim or int
        :param heads: number of parallel de/rasterizations, int > 0
        :param dim: dimension of the feature tensor to de/rasterize, 2 or 3
        '''
        super().__init__()
        self.dim = dim
        self.heads = heads
        if isinstance(tensor_size, int):
            self.tensor_size = dim * [tensor_size]
        else:
            assert isinstance(tensor_size, tuple)
            assert len(tensor_size) == dim
            self.tensor_size = tensor_size
        tensor_mod = torch.tensor(self.tensor_size, dtype=torch.float32, requires_grad=False)[None, :, None]
        self.register_buffer('tensor_mod', tensor_mod)
        self.spread_size = 8 if self.dim == 3 else 4
        self.eps = 1e-7
        self._geom = CF.Geometry(self.tensor_size, self.heads, self.dim)


def _handle_of(local_coordinate, flattened_index):
    """The PositionsHandle if both tensors are the untouched outputs of one DifferentiablePositions call."""
    if not CF.config.fused:
        return None
    tag = getattr(local_coordinate, "_ctb_tag", None)
    if tag is None or getattr(flattened_index, "_ctb_tag", None) is not tag:
        return None
    handle, lc_version, idx_version, keys_version = tag
    if local_coordinate._version != lc_version or flattened_index._version != idx_version:
        return None  # modified in place since we produced them
    if handle.keys._version != keys_version:
        return None  # the keys were modified in place: lc / idx still describe the old positions, use them as given
    return handle


class DifferentiablePositions(DifferentiableGridModule):
    '''Bi/tri-linear coordinates of each point w.r.t. the enclosing feature-map cell.'''

    def forward(self, keys):
        '''
        :param keys: float32 [batch_size, heads * dim, num_points] in (-1, 1)
        :return: local_coordinate float32 [batch_size, heads, 2^dim, num_points],
                 flattened_index  int64   [batch_size, heads, 2^dim, num_points]
        '''
        assert keys.size(1) == self.heads * self.dim
        if keys.numel() == 0:
            # empty batch / empty cloud: the reference's elementwise ops return empty tensors of these shapes
            B, N, S = keys.size(0), keys.size(2), 2 ** self.dim
            return (keys.new_zeros((B, self.heads, S, N)) + 0.0 * keys.sum(),
                    torch.zeros((B, self.heads, S, N), dtype=torch.int64, device=keys.device))
        local_coordinate, flattened_index = CF.positions(keys, self._geom)
        tag = (CF.PositionsHandle(keys, self._geom), local_coordinate._version, flattened_index._version, keys._version)
        local_coordinate._ctb_tag = tag
        flattened_index._ctb_tag = tag
        return local_coordinate, flattened_index


class Splat(DifferentiableGridModule):
    '''Differentiable rasterization (splatting) into a 2D/3D feature grid.

    reduce="max" is the reference (scatter_max onto a zero grid); reduce="sum" is the scatter-add variant.'''

    def __init__(self, tensor_size=20, heads=4, dim=3, reduce="max", out_dtype=None):
        super().__init__(tensor_size, heads, dim)
        assert reduce in ("max", "sum")
        self.reduce = reduce
        self.out_dtype = out_dtype      # torch.bfloat16: bf16 grid storage mode (extension; reference returns fp32)

    def forward(self, local_coordinate, flattened_index, features, pts_padding=None):
        '''
        :return: feature map [batch_size, heads * feature_dim, *tensor_size]
        '''
        assert features.dtype == torch.float32
        assert features.size(1) % self.heads == 0
        if features.numel() == 0:
            # nothing to splat: scatter_max leaves its zero-initialised output untouched (cloud_transform.py:164-173)
            z = features.new_zeros((features.size(0), features.size(1)) + tuple(self._geom.sizes)) + 0.0 * features.sum()
            return z if self.out_dtype is None else z.to(self.out_dtype)
        reduce = _lib.REDUCE_MAX if self.reduce == "max" else _lib.REDUCE_SUM
        handle = _handle_of(local_coordinate, flattened_index)
        if handle is not None and handle.geom.sizes == self._geom.sizes and handle.geom.heads == self.heads:
            return CF.fused_splat(handle, features, pts_padding, reduce, self.out_dtype)
        z = CF.splat(local_coordinate, flattened_index, features, pts_padding, self._geom, reduce)
        return z if self.out_dtype is None else z.to(self.out_dtype)


class Slice(DifferentiableGridModule):
    '''Differentiable sampling of a 2D/3D feature grid.'''

    def forward(self, local_coordinate, flattened_index, convolved, pts_padding=None):
        '''
        :return: sliced features float32 [batch_size, heads * feature_dim, num_points]
        '''
        assert convolved.size(1) % self.heads == 0
        if local_coordinate.numel() == 0 or convolved.numel() == 0:
            # no points: torch.gather on an empty index returns an empty tensor (cloud_transform.py:216-221)
            out = convolved.new_zeros((convolved.size(0), convolved.size(1), local_coordinate.size(-1)), dtype=torch.float32)
            return out + 0.0 * convolved.float().sum()
        handle = _handle_of(local_coordinate, flattened_index)
        if handle is not None and handle.geom.sizes == self._geom.sizes and handle.geom.heads == self.heads:
            return CF.fused_slice(handle, convolved, pts_padding)
        return CF.slice_(local_coordinate, flattened_index, convolved, pts_padding, self._geom)
