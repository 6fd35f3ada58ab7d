"""Mint model-level golden fixtures by running the UNMODIFIED reference files (under the two dependency shims of
oracle/reference_loader.py) on CPU with seeded weights and inputs.  Build-container only; the .npz files are committed.

    python tests/golden/make_model_golden.py

  classifier_golden.npz   model_zoo/scanobject/classifier.py:35-148 (24.02 M parameters, weights = torch.manual_seed(0)
                          initialisation, reproduced by the test with the same seed): input cloud, class / mask
                          predictions, lattice statistics, d loss / d input and two parameter gradients.
  blocks_golden.npz       layers/multihead_ct.py:9-118 MultiHead (2-D, 3-D with scales, 2-D with a padding mask),
                          layers/multihead_ct_pool.py:9-86 MultiHeadPool, layers/multihead_ct.py:121-198 MultiHeadUnion
                          and layers/multihead_ct_adain.py:19-218 MultiHeadAdaIn / MultiHeadUnionAdaIn with
                          non-trivial parameters: state_dict + inputs -> outputs, stats and gradients.  Pins the
                          mirrors of cloud_transformers_b200/mhct.py (rows A8 / N1) to the reference's own blocks.
Dropout layers are put in eval mode on both sides (their masks come from device-specific RNG streams); BatchNorm
stays in training mode.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_loader as RL  # noqa: E402


def dropout_off(model):
    for m in model.modules():
        if "Dropout" in type(m).__name__:
            m.eval()


def surface_cloud(gen, B, N):
    """points on a few random planes / spheres, centred and max-norm normalised (datasets/scanobjectnn.py:44-62)"""
    k = 3
    which = torch.randint(0, k, (B, N), generator=gen)
    u = torch.rand(B, N, 2, generator=gen) * 2 - 1
    basis = torch.randn(B, k, 3, 3, generator=gen)
    offs = torch.randn(B, k, 3, generator=gen) * 0.3
    b = torch.gather(basis, 1, which[:, :, None, None].expand(-1, -1, 3, 3))
    o = torch.gather(offs, 1, which[:, :, None].expand(-1, -1, 3))
    pts = u[..., 0:1] * b[:, :, 0] + u[..., 1:2] * b[:, :, 1] + o
    pts = pts - pts.mean(1, keepdim=True)
    pts = pts / pts.norm(dim=-1).max(dim=1)[0][:, None, None]
    return pts.permute(0, 2, 1).contiguous()


def randomize(model, gen, scale=0.3):
    """make zero-initialised / identity parameters non-trivial so that every path carries signal"""
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("key_bn.weight") or name == "scale" or name.endswith(".scale"):
                p.copy_(torch.randn(p.shape, generator=gen) * scale if p.dim() else torch.tensor(0.37))
            if name.endswith("shift") or name.endswith("scales"):
                p.add_(torch.randn(p.shape, generator=gen) * 0.1)


def stats_to_np(stats):
    out = []
    for st in stats:
        out.append([float(st[0]), float(st[1]), float(st[2])])
    return np.asarray(out, dtype=np.float64)


def mint_classifier(out_path):
    gen = torch.Generator().manual_seed(123)
    B, N = 2, 256
    pcd = surface_cloud(gen, B, N)[:, :, None, :].contiguous()          # [B, 3, 1, N] as train_classification.py:195
    cw = torch.randn(B, 15, generator=gen)
    mw = torch.randn(B, 1, 1, N, generator=gen)
    with RL.reference_tree(dropin=False) as rt:
        torch.manual_seed(0)
        model = rt.load_model("model_zoo/scanobject/classifier.py")
        model.train()
        dropout_off(model)
        x = pcd.clone().requires_grad_(True)
        class_pred, mask_pred, stats = model(x)
        loss = (class_pred * cw).sum() + (mask_pred * mw).sum()
        loss.backward()
        params = dict(model.named_parameters())
        np.savez_compressed(
            out_path, pcd=pcd.numpy(), cw=cw.numpy(), mw=mw.numpy(), class_pred=class_pred.detach().numpy(),
            mask_pred=mask_pred.detach().numpy(), stats=stats_to_np(stats), grad_pcd=x.grad.numpy(),
            grad_first_conv=params["first_process.0.weight"].grad.numpy(),
            grad_shift0=params["attentions_encoder.0.attentions.0.transform.shift"].grad.numpy(),
            grad_kv11=params["attentions_encoder.11.attentions.1.keys_values_pred.0.weight"].grad.numpy()[:8],
            n_params=np.asarray(sum(p.numel() for p in model.parameters())))
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


def mint_blocks(out_path):
    out = {}
    gen = torch.Generator().manual_seed(77)
    with RL.reference_tree(dropin=False):
        import layers.multihead_ct as mh
        import layers.multihead_ct_adain as mha
        import layers.multihead_ct_pool as mhp

        def save_state(prefix, module):
            for k, v in module.state_dict().items():
                out["%s/sd/%s" % (prefix, k)] = v.numpy().copy()

        def run(prefix, module, x, pcd, style=None, pad=None, returns_stats_list=False):
            module.train()
            save_state(prefix, module)
            xi = x.clone().requires_grad_(True)
            pi = pcd.clone().requires_grad_(True)
            args = (xi, (pi, pad)) if pad is not None else (xi, pi)
            if style is not None:
                si = style.clone().requires_grad_(True)
                res, stats = module(xi, si, pi)          # (input, style, orig_pcd), multihead_ct_adain.py:107
            else:
                res, stats = module(*args)
            gw = torch.randn(res.shape, generator=gen)
            (res * gw).sum().backward()
            out[prefix + "/x"] = x.numpy()
            out[prefix + "/pcd"] = pcd.numpy()
            out[prefix + "/gw"] = gw.numpy()
            out[prefix + "/out"] = res.detach().numpy()
            out[prefix + "/gx"] = xi.grad.numpy()
            out[prefix + "/gpcd"] = pi.grad.numpy()
            if pad is not None:
                out[prefix + "/pad"] = pad.numpy()
            if style is not None:
                out[prefix + "/style"] = style.numpy()
                out[prefix + "/gstyle"] = si.grad.numpy()
            sl = stats if returns_stats_list else [stats]
            out[prefix + "/stats"] = np.asarray([[float(s[0]), float(s[1]), float(s[2])] for s in sl], dtype=np.float64)
            for k, p in module.named_parameters():
                if p.grad is not None and (k.endswith("log_R") or k.endswith("shift") or k.endswith("scales") or
                                           k.endswith("key_bn.weight") or k == "scale" or k.endswith(".scale") or
                                           k.endswith("conv.0.bias")):
                    out["%s/gp/%s" % (prefix, k)] = p.grad.numpy().copy()

        B, N, D = 2, 200, 32
        x = torch.randn(B, D, N, generator=gen)
        pcd = surface_cloud(gen, B, N)
        torch.manual_seed(5)
        m = mh.MultiHead(model_dim=D, in_feature_dim=4, out_model_dim=D, tensor_size=16, tensor_dim=2, heads=4)
        randomize(m, gen)
        run("mh2d", m, x, pcd)
        m = mh.MultiHead(model_dim=D, in_feature_dim=8, out_model_dim=D, tensor_size=8, tensor_dim=3, heads=4, scales=True)
        randomize(m, gen)
        run("mh3d_scales", m, x, pcd)
        m = mh.MultiHead(model_dim=D, in_feature_dim=4, out_model_dim=D, tensor_size=12, tensor_dim=2, heads=2)
        randomize(m, gen)
        run("mh2d_pad", m, x, pcd, pad=(torch.rand(B, N, generator=gen) > 0.3).float())
        m = mhp.MultiHeadPool(model_dim=D, in_feature_dim=8, tensor_size=8, tensor_dim=3, heads=4)
        randomize(m, gen)
        run("pool3d", m, x, pcd)
        m = mh.MultiHeadUnion(model_dim=D, features_dims=[4, 8], heads=[4, 4], tensor_sizes=[16, 8], model_dim_out=D,
                              tensor_dims=[2, 3])
        randomize(m, gen)
        run("union", m, x, pcd, returns_stats_list=True)
        style = torch.randn(B, 24, generator=gen)
        m = mha.MultiHeadAdaIn(model_dim=D, in_feature_dim=4, out_model_dim=D, tensor_size=16, tensor_dim=2, heads=4,
                               n_latent=24)
        randomize(m, gen)
        run("adain2d", m, x, pcd, style=style)
        m = mha.MultiHeadUnionAdaIn(model_dim=D, features_dims=[4, 8], heads=[4, 4], tensor_sizes=[16, 8],
                                    model_dim_out=D, tensor_dims=[2, 3], n_latent=24)
        randomize(m, gen)
        run("union_adain", m, x, pcd, style=style, returns_stats_list=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    mint_blocks(os.path.join(HERE, "blocks_golden.npz"))
    mint_classifier(os.path.join(HERE, "classifier_golden.npz"))
