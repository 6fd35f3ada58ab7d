"""Host-side mirror of the reference's MHCT blocks (layers/multihead_ct.py:9-198, layers/multihead_ct_pool.py:9-86,
layers/multihead_ct_adain.py:8-218)
wired to the fused B200 kernels -- the callers either side of the hot path.

Same constructor arguments, sub-module / parameter names (so reference checkpoints load with strict=True) and
return contract `(result, stats)` as the reference; the forward differs only in HOW the path is computed:
  * keys: one fused kernel for shift + rotation + scales + tanh (ctb_project_fwd) instead of add / einsum / tanh
  * positions are never materialised: Splat / Slice take the keys directly (fused_splat / fused_slice)
  * the occupancy statistic is one counting pass without host sync (ctb_count_occupied)
The dense contractions (1x1 Conv1d, BatchNorm, grouped 3x3(x3) conv) stay in PyTorch / cuDNN, as the north star says.
"""
import torch
from torch import nn

from . import functional as CF
from .cloud_transform import DifferentiablePositions, Slice, Splat
from .so3 import so3_exponential_map


class _Transformer(nn.Module):
    """Parameters of VolTransformer / PlaneTransformer (layers/utils.py:9-61): log_R, shift, optional scales."""

    def __init__(self, heads, dim, scales=False):
        super().__init__()
        self.heads, self.dim = heads, dim
        self.log_R = nn.Parameter(torch.randn(heads, 3, dtype=torch.float32))
        self.shift = nn.Parameter(torch.zeros(heads, 3, dtype=torch.float32))
        self.do_scales = scales
        if scales:
            self.scales = nn.Parameter(torch.ones(heads, dim, dtype=torch.float32))

    def keys(self, orig_pcd, keys_res, res_scale=None):
        """-> (lattice = tanh(keys) [B, heads*dim, N], (mean, var) of the pre-tanh keys as 0-dim device tensors)"""
        rot = so3_exponential_map(self.log_R)
        acc = torch.zeros(2, dtype=torch.float64, device=orig_pcd.device)
        lattice = CF.project_keys(orig_pcd, keys_res, self.shift, rot, self.scales if self.do_scales else None,
                                  res_scale, heads=self.heads, dim=self.dim, key_stats=acc)
        return lattice, CF.key_mean_var(acc, lattice.numel())


class MultiHead(nn.Module):
    def __init__(self, model_dim, in_feature_dim, out_model_dim, tensor_size, tensor_dim, heads, scales=False):
        super().__init__()
        assert tensor_dim in (2, 3)
        self.in_feature_dim, self.out_model_dim = in_feature_dim, out_model_dim
        self.model_dim, self.tensor_size, self.tensor_dim, self.heads = model_dim, tensor_size, tensor_dim, heads
        self.keys_values_pred = nn.Sequential(nn.Conv1d(model_dim, heads * (in_feature_dim + 3), kernel_size=1, bias=False))
        self.values_bn = nn.BatchNorm1d(heads * in_feature_dim)
        self.key_bn = nn.BatchNorm1d(heads * 3)
        # kept for their `tensor_mod` buffers (checkpoint compatibility) and for callers that want lc / idx
        self.diff_poss = DifferentiablePositions(tensor_size=tensor_size, dim=tensor_dim, heads=heads)
        self.splat = Splat(tensor_size=tensor_size, dim=tensor_dim, heads=heads)
        self.slice = Slice(tensor_size=tensor_size, dim=tensor_dim, heads=heads)
        conv = nn.Conv3d if tensor_dim == 3 else nn.Conv2d
        self.conv = nn.Sequential(conv(heads * in_feature_dim, heads * in_feature_dim, kernel_size=3, stride=1, padding=1,
                                       groups=heads, bias=True))
        self.after = nn.Sequential(nn.BatchNorm1d(heads * in_feature_dim), nn.ReLU(inplace=True))
        self.transform = _Transformer(heads, tensor_dim, scales)
        sizes = [tensor_size] * tensor_dim if isinstance(tensor_size, int) else list(tensor_size)
        self._geom = CF.Geometry(sizes, heads, tensor_dim)
        torch.nn.init.zeros_(self.key_bn.weight)       # multihead_ct.py:79-80

    def forward(self, input, orig_pcd, return_lattice=False):
        if isinstance(orig_pcd, tuple):                # (points, padding mask), multihead_ct.py:84-87
            orig_pcd, pts_padd = orig_pcd
        else:
            pts_padd = None
        key_values = self.keys_values_pred(input)
        keys_res = self.key_bn(key_values[:, :self.heads * 3])
        values = self.values_bn(key_values[:, self.heads * 3:])
        lattice, (k_mean, k_var) = self.transform.keys(orig_pcd, keys_res)  # A8, fused; exact pre-tanh key statistics
        handle = CF.PositionsHandle(lattice, self._geom)
        z = CF.fused_splat(handle, values, pts_padd)                        # A1 + A2 + A3
        with torch.no_grad():
            occ = CF.count_occupied(z).float() / (input.size(0) * self.in_feature_dim * self.heads)   # A9
        result = self.after(CF.fused_slice(handle, self.conv(z), pts_padd))  # A4
        stats = (occ, k_mean, k_var, None)                                  # multihead_ct.py:109-113
        if return_lattice:
            result = result, lattice
        return result, stats


class MultiHeadPool(nn.Module):
    def __init__(self, model_dim, in_feature_dim, tensor_size, tensor_dim, heads, scales=False):
        super().__init__()
        assert tensor_dim in (2, 3)
        self.in_feature_dim, self.model_dim = in_feature_dim, model_dim
        self.tensor_size, self.tensor_dim, self.heads = tensor_size, tensor_dim, heads
        self.keys_values_pred = nn.Sequential(nn.Conv1d(model_dim, heads * (in_feature_dim + 3), kernel_size=1, bias=False))
        self.values_bn = nn.BatchNorm1d(heads * in_feature_dim)
        self.key_bn = nn.BatchNorm1d(heads * 3)
        self.diff_poss = DifferentiablePositions(tensor_size=tensor_size, dim=tensor_dim, heads=heads)
        self.splat = Splat(tensor_size=tensor_size, dim=tensor_dim, heads=heads)
        self.transform = _Transformer(heads, tensor_dim, scales)
        sizes = [tensor_size] * tensor_dim if isinstance(tensor_size, int) else list(tensor_size)
        self._geom = CF.Geometry(sizes, heads, tensor_dim)
        torch.nn.init.zeros_(self.key_bn.weight)

    def forward(self, input, orig_pcd, return_lattice=False):
        key_values = self.keys_values_pred(input)
        keys_res = self.key_bn(key_values[:, :self.heads * 3])
        values = self.values_bn(key_values[:, self.heads * 3:])
        lattice, (k_mean, k_var) = self.transform.keys(orig_pcd, keys_res)
        z = CF.fused_splat(CF.PositionsHandle(lattice, self._geom), values)
        with torch.no_grad():
            occ = CF.count_occupied(z).float() / (input.size(0) * self.in_feature_dim * self.heads)
        stats = (occ, k_mean, k_var, None)
        result = (z, lattice) if return_lattice else z
        return result, stats


class MultiHeadUnion(nn.Module):
    def __init__(self, model_dim, features_dims, tensor_sizes, tensor_dims, heads, model_dim_out=None, scales=False):
        super().__init__()
        assert len(features_dims) == len(tensor_sizes) == len(tensor_dims) == len(heads)
        self.model_dim = model_dim
        self.model_dim_out = model_dim if model_dim_out is None else model_dim_out
        self.prenorm = nn.Sequential()
        self.after = nn.Sequential(
            nn.Conv1d(sum(h * f for h, f in zip(heads, features_dims)), self.model_dim_out, kernel_size=1, bias=False),
            nn.BatchNorm1d(self.model_dim_out), nn.ReLU(inplace=True))
        self.shortcut = nn.Sequential()
        if self.model_dim != self.model_dim_out:
            self.shortcut.add_module('shortcut_conv', nn.Conv1d(self.model_dim, self.model_dim_out, kernel_size=1, bias=False))
            self.shortcut.add_module('shortcut_bn', nn.BatchNorm1d(self.model_dim_out))
        self.attentions = nn.ModuleList([
            MultiHead(model_dim=model_dim, in_feature_dim=f, out_model_dim=self.model_dim_out, tensor_size=t, tensor_dim=d,
                      heads=h, scales=scales) for f, t, d, h in zip(features_dims, tensor_sizes, tensor_dims, heads)])

    def forward(self, x, orig_pcd):
        x = self.prenorm(x)
        residual = self.shortcut(x)
        results, stats = [], []
        for attention in self.attentions:
            r, s = attention(x, orig_pcd)
            results.append(r)
            stats.append(s)
        return residual + self.after(torch.cat(results, dim=1)), stats


class AdaIn1dUpd(nn.Module):
    """layers/utils.py:82-97: InstanceNorm1d followed by a style-conditioned affine map."""

    def __init__(self, num_features, num_latent):
        super().__init__()
        self.num_features, self.num_latent = num_features, num_latent
        self.instance_norm = nn.InstanceNorm1d(num_features, eps=1e-5, affine=False)
        self.linear = nn.Linear(num_latent, num_features * 2)

    def forward(self, x, z):
        x = self.instance_norm(x)
        var_bias = self.linear(z).reshape(-1, 2, self.num_features)
        return x * (var_bias[:, 0][:, :, None] + 1) + var_bias[:, 1][:, :, None]


def forward_style(module_list, input, z):
    """layers/multihead_ct_adain.py:8-16"""
    for layer in module_list:
        input = layer(input, z) if isinstance(layer, AdaIn1dUpd) else layer(input)
    return input


class MultiHeadAdaIn(nn.Module):
    """layers/multihead_ct_adain.py:19-136 (the decoder block of the completion / reconstruction models): AdaIN in
    place of BatchNorm, a learnable scalar `scale` on the key residual (folded into the fused projection kernel as
    res_scale), no padding mask.  The reference copies occupancy / key statistics AND the full keys tensor to the
    host on every forward (:127-131); here the statistics stay on the device and the keys are not copied."""

    def __init__(self, model_dim, in_feature_dim, out_model_dim, tensor_size, tensor_dim, heads, n_latent=256, unet=False,
                 scales=False):
        super().__init__()
        assert tensor_dim in (2, 3)
        self.in_feature_dim, self.out_model_dim = in_feature_dim, out_model_dim
        self.model_dim, self.tensor_size, self.tensor_dim, self.heads = model_dim, tensor_size, tensor_dim, heads
        self.num_latent = n_latent
        self.keys_values_pred = nn.Sequential(nn.Conv1d(model_dim, heads * (in_feature_dim + 3), kernel_size=1, bias=False))
        self.values_bn = nn.Sequential(AdaIn1dUpd(heads * in_feature_dim, num_latent=n_latent))
        self.keys_bn = nn.Sequential(AdaIn1dUpd(heads * 3, num_latent=n_latent))
        self.diff_poss = DifferentiablePositions(tensor_size=tensor_size, dim=tensor_dim, heads=heads)
        self.splat = Splat(tensor_size=tensor_size, dim=tensor_dim, heads=heads)
        self.slice = Slice(tensor_size=tensor_size, dim=tensor_dim, heads=heads)
        conv = nn.Conv3d if tensor_dim == 3 else nn.Conv2d
        self.conv = nn.Sequential(conv(heads * in_feature_dim, heads * in_feature_dim, kernel_size=3, stride=1, padding=1,
                                       groups=heads, bias=True))
        self.after = nn.Sequential(AdaIn1dUpd(heads * in_feature_dim, num_latent=n_latent), nn.ReLU(inplace=True))
        self.scale = nn.Parameter(data=torch.tensor(0, dtype=torch.float32), requires_grad=True)
        self.transform = _Transformer(heads, tensor_dim, scales)
        sizes = [tensor_size] * tensor_dim if isinstance(tensor_size, int) else list(tensor_size)
        self._geom = CF.Geometry(sizes, heads, tensor_dim)

    def forward(self, input, style, orig_pcd, return_lattice=False):
        key_values = forward_style(self.keys_values_pred, input, style)
        keys_res = forward_style(self.keys_bn, key_values[:, :self.heads * 3], style)
        values = forward_style(self.values_bn, key_values[:, self.heads * 3:], style)
        lattice, (k_mean, k_var) = self.transform.keys(orig_pcd, keys_res, res_scale=self.scale)
        handle = CF.PositionsHandle(lattice, self._geom)
        z = CF.fused_splat(handle, values)
        with torch.no_grad():
            occ = CF.count_occupied(z).float() / (input.size(0) * self.in_feature_dim * self.heads)
        result = forward_style(self.after, CF.fused_slice(handle, self.conv(z)), style)
        stats = (occ, k_mean, k_var, None)
        if return_lattice:
            result = result, lattice
        return result, stats


class MultiHeadUnionAdaIn(nn.Module):
    """layers/multihead_ct_adain.py:139-218"""

    def __init__(self, model_dim, features_dims, tensor_sizes, tensor_dims, heads, model_dim_out=None, n_latent=256,
                 unet=False, scales=False):
        super().__init__()
        assert len(features_dims) == len(tensor_sizes) == len(tensor_dims) == len(heads)
        self.model_dim = model_dim
        self.model_dim_out = model_dim if model_dim_out is None else model_dim_out
        self.prenorm = nn.Sequential()
        self.after = nn.Sequential(
            nn.Conv1d(sum(h * f for h, f in zip(heads, features_dims)), self.model_dim_out, kernel_size=1, bias=False),
            AdaIn1dUpd(self.model_dim_out, num_latent=n_latent), nn.ReLU(inplace=True))
        self.shortcut = nn.Sequential()
        if self.model_dim != self.model_dim_out:
            self.shortcut.add_module('shortcut_conv', nn.Conv1d(self.model_dim, self.model_dim_out, kernel_size=1, bias=False))
            self.shortcut.add_module('shortcut_bn', AdaIn1dUpd(self.model_dim_out, num_latent=n_latent))
        self.attentions = nn.ModuleList([
            MultiHeadAdaIn(model_dim=model_dim, in_feature_dim=f, out_model_dim=self.model_dim_out, tensor_size=t,
                           tensor_dim=d, n_latent=n_latent, heads=h, unet=unet, scales=scales)
            for f, t, d, h in zip(features_dims, tensor_sizes, tensor_dims, heads)])

    def forward(self, x, style, orig_pcd):
        x = self.prenorm(x)
        residual = forward_style(self.shortcut, x, style)
        results, stats = [], []
        for attention in self.attentions:
            r, s = attention(x, style, orig_pcd)
            results.append(r)
            stats.append(s)
        return residual + forward_style(self.after, torch.cat(results, dim=1), style), stats


class ScanObjectTrunk(nn.Module):
    """The MHCT trunk of the reference's ScanObjectNN classifier (model_zoo/scanobject/classifier.py:41-85, :114-131):
    Conv1d(3 -> 512) + BN + ReLU, 12 MultiHeadUnion blocks (4 x {2D 128^2 F4 + 3D 32^3 F4, 2D 64^2 F16 + 3D 16^3 F16,
    2D 16^2 F16 + 3D 8^3 F32}, 16 heads each) and the two MultiHeadPool rasterisations (3D 8^3 F32, 2D 16^2 F16).
    The Res2D/Res3D towers behind the pools (dense cuDNN convolutions, out of scope here) are replaced by a global
    average pool + Linear so the trunk can be trained end to end; bench.py uses it for the MHCT training throughput."""

    def __init__(self, n_classes=15, model_dim=512, n_rounds=4):
        super().__init__()
        self.model_dim = model_dim
        self.first_process = nn.Sequential(nn.Conv1d(3, model_dim, kernel_size=1, bias=False), nn.BatchNorm1d(model_dim),
                                           nn.ReLU(inplace=True))
        blocks = []
        for _ in range(n_rounds):
            for feats, sizes in (([4, 4], [128, 32]), ([16, 16], [64, 16]), ([16, 32], [16, 8])):
                blocks.append(MultiHeadUnion(model_dim=model_dim, features_dims=feats, heads=[16, 16], tensor_sizes=sizes,
                                             model_dim_out=model_dim, tensor_dims=[2, 3]))
        self.attentions_encoder = nn.ModuleList(blocks)
        self.pool3d = MultiHeadPool(model_dim=model_dim, in_feature_dim=32, heads=16, tensor_size=8, tensor_dim=3)
        self.pool2d = MultiHeadPool(model_dim=model_dim, in_feature_dim=16, heads=16, tensor_size=16, tensor_dim=2)
        self.class_vector = nn.Sequential(nn.Linear(32 * 16 + 16 * 16, 1024), nn.BatchNorm1d(1024), nn.ReLU(inplace=True))
        self.class_head = nn.Sequential(nn.Dropout(0.5), nn.Linear(1024, n_classes))

    def forward(self, pcd):
        """pcd [B, 3, N] -> class logits [B, n_classes], list of per-block lattice statistics."""
        x = self.first_process(pcd)
        stats = []
        for block in self.attentions_encoder:
            x, st = block(x, pcd)
            stats += st
        to_3d, st3 = self.pool3d(x, pcd)
        to_2d, st2 = self.pool2d(x, pcd)
        stats += [st3, st2]
        pooled = torch.cat([to_3d.flatten(2).mean(-1), to_2d.flatten(2).mean(-1)], dim=-1)
        return self.class_head(self.class_vector(pooled)), stats
