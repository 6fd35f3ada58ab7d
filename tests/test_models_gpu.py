"""Model-level parity (-m gpu): the reference's OWN model / block files running on the B200 kernels through the drop-in
(`dropin/layers/cloud_transform.py` ahead of the reference tree on sys.path, nothing else changed), and the block
mirrors of cloud_transformers_b200/mhct.py, both against fixtures minted on CPU from the unmodified reference
(tests/golden/make_model_golden.py).

The reference tree comes from oracle/_ref (verbatim copy staged by oracle/stage_ref.py in the build container; it
travels to the GPU box) -- test infrastructure only.  Tolerance: fp32 end to end (TF32 off), rel 2e-3 of the tensor's
magnitude: CPU and GPU differ in the summation order of every BatchNorm / convolution in front of the keys, and a key
that moves by one ulp across a cell boundary changes which cell a point feeds.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import reference_loader as RL  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _fp32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _need_tree():
    if not RL.available():
        pytest.skip("reference tree not staged (oracle/_ref): run `python oracle/stage_ref.py` in the build container")


def close(a, b, what, rel=2e-3):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(float(np.abs(b).max()), 1e-12)
    err = float(np.abs(a - b).max()) / scale
    assert err <= rel, "%s: max |diff| / max |ref| = %.3e > %.1e" % (what, err, rel)
    return err


def close_l2(a, b, what, rel):
    """for gradients through the whole 12-block network: a handful of arg-max winners flip between CPU and GPU, which
    moves individual entries by O(1e-2) of the maximum; the relative L2 error bounds the aggregate."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    assert err <= rel, "%s: relative L2 error %.3e > %.1e" % (what, err, rel)


def dropout_off(model):
    for m in model.modules():
        if "Dropout" in type(m).__name__:
            m.eval()


def test_reference_classifier_runs_unchanged_through_dropin():
    """model_zoo/scanobject/classifier.py:35-148, exec'd like utils/train_util.py:23-27, 24.02 M parameters."""
    _need_tree()
    g = np.load(os.path.join(GOLD, "classifier_golden.npz"))
    with RL.reference_tree(dropin=True) as rt:
        torch.manual_seed(0)
        model = rt.load_model("model_zoo/scanobject/classifier.py")
        import layers.cloud_transform as ct
        import layers.multihead_ct as mh
        assert os.path.realpath(ct.__file__).startswith(os.path.realpath(RL.DROPIN_ROOT))
        assert os.path.realpath(mh.__file__).startswith(os.path.realpath(rt.root))
        assert type(model.pool3d.splat).__module__ == "cloud_transformers_b200.cloud_transform"
        assert sum(p.numel() for p in model.parameters()) == int(g["n_params"]) == 24021136
        model = model.to(DEV).train()
        dropout_off(model)
        x = torch.from_numpy(g["pcd"]).to(DEV).requires_grad_(True)
        class_pred, mask_pred, stats = model(x)
        loss = (class_pred * torch.from_numpy(g["cw"]).to(DEV)).sum() + (mask_pred * torch.from_numpy(g["mw"]).to(DEV)).sum()
        loss.backward()
        torch.cuda.synchronize()
        params = dict(model.named_parameters())
        close(class_pred.detach().cpu().numpy(), g["class_pred"], "class_pred")
        close(mask_pred.detach().cpu().numpy(), g["mask_pred"], "mask_pred")
        st = np.asarray([[float(s[0]), float(s[1]), float(s[2])] for s in stats])
        close(st, g["stats"], "lattice statistics")
        close_l2(x.grad.cpu().numpy(), g["grad_pcd"], "d loss / d input cloud", rel=2e-2)
        close_l2(params["first_process.0.weight"].grad.cpu().numpy(), g["grad_first_conv"], "grad first conv", rel=2e-2)
        close_l2(params["attentions_encoder.0.attentions.0.transform.shift"].grad.cpu().numpy(), g["grad_shift0"],
                 "grad shift of block 0", rel=2e-2)
        close_l2(params["attentions_encoder.11.attentions.1.keys_values_pred.0.weight"].grad.cpu().numpy()[:8],
                 g["grad_kv11"], "grad keys/values projection of block 11", rel=2e-2)


BLOCKS = {
    # prefix: (class name, ctor kwargs, takes style, list of stats)
    "mh2d": ("MultiHead", dict(model_dim=32, in_feature_dim=4, out_model_dim=32, tensor_size=16, tensor_dim=2, heads=4)),
    "mh3d_scales": ("MultiHead", dict(model_dim=32, in_feature_dim=8, out_model_dim=32, tensor_size=8, tensor_dim=3, heads=4,
                                      scales=True)),
    "mh2d_pad": ("MultiHead", dict(model_dim=32, in_feature_dim=4, out_model_dim=32, tensor_size=12, tensor_dim=2, heads=2)),
    "pool3d": ("MultiHeadPool", dict(model_dim=32, in_feature_dim=8, tensor_size=8, tensor_dim=3, heads=4)),
    "union": ("MultiHeadUnion", dict(model_dim=32, features_dims=[4, 8], heads=[4, 4], tensor_sizes=[16, 8], model_dim_out=32,
                                     tensor_dims=[2, 3])),
    "adain2d": ("MultiHeadAdaIn", dict(model_dim=32, in_feature_dim=4, out_model_dim=32, tensor_size=16, tensor_dim=2, heads=4,
                                       n_latent=24)),
    "union_adain": ("MultiHeadUnionAdaIn", dict(model_dim=32, features_dims=[4, 8], heads=[4, 4], tensor_sizes=[16, 8],
                                                model_dim_out=32, tensor_dims=[2, 3], n_latent=24)),
}


def _run_block(module, g, prefix):
    module = module.to(DEV).train()
    x = torch.from_numpy(g[prefix + "/x"]).to(DEV).requires_grad_(True)
    pcd = torch.from_numpy(g[prefix + "/pcd"]).to(DEV).requires_grad_(True)
    if prefix + "/style" in g.files:
        style = torch.from_numpy(g[prefix + "/style"]).to(DEV).requires_grad_(True)
        res, stats = module(x, style, pcd)
    elif prefix + "/pad" in g.files:
        style = None
        res, stats = module(x, (pcd, torch.from_numpy(g[prefix + "/pad"]).to(DEV)))
    else:
        style = None
        res, stats = module(x, pcd)
    (res * torch.from_numpy(g[prefix + "/gw"]).to(DEV)).sum().backward()
    torch.cuda.synchronize()
    close(res.detach().cpu().numpy(), g[prefix + "/out"], prefix + " output")
    close(x.grad.cpu().numpy(), g[prefix + "/gx"], prefix + " grad input", rel=5e-3)
    close(pcd.grad.cpu().numpy(), g[prefix + "/gpcd"], prefix + " grad cloud", rel=5e-3)
    if style is not None:
        close(style.grad.cpu().numpy(), g[prefix + "/gstyle"], prefix + " grad style", rel=5e-3)
    sl = stats if isinstance(stats, list) else [stats]
    close(np.asarray([[float(s[0]), float(s[1]), float(s[2])] for s in sl]), g[prefix + "/stats"], prefix + " stats")
    named = dict(module.named_parameters())
    for k in [k for k in g.files if k.startswith(prefix + "/gp/")]:
        if k.endswith("conv.0.bias"):
            continue     # the normalisation behind the Slice cancels a per-channel bias: this gradient is rounding noise
        close(named[k.split("/gp/", 1)[1]].grad.cpu().numpy(), g[k], k, rel=5e-3)


@pytest.mark.parametrize("prefix", sorted(BLOCKS))
def test_block_mirror_matches_reference_fixture(prefix):
    """cloud_transformers_b200/mhct.py (fused projection + tanh, fused Splat / Slice, device-side statistics) loads the
    reference block's state_dict strictly and reproduces its outputs, statistics and gradients."""
    from cloud_transformers_b200 import mhct
    g = np.load(os.path.join(GOLD, "blocks_golden.npz"))
    cls, kw = BLOCKS[prefix]
    module = getattr(mhct, cls)(**kw)
    sd = {k.split("/sd/", 1)[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix + "/sd/")}
    module.load_state_dict(sd, strict=True)
    _run_block(module, g, prefix)


@pytest.mark.parametrize("prefix", sorted(BLOCKS))
def test_reference_block_through_dropin_matches_fixture(prefix):
    """the reference's own layers/multihead_ct*.py on top of dropin/layers/cloud_transform.py"""
    _need_tree()
    g = np.load(os.path.join(GOLD, "blocks_golden.npz"))
    cls, kw = BLOCKS[prefix]
    with RL.reference_tree(dropin=True):
        import layers.multihead_ct as mh
        import layers.multihead_ct_adain as mha
        import layers.multihead_ct_pool as mhp
        ctor = getattr(mh, cls, None) or getattr(mha, cls, None) or getattr(mhp, cls)
        module = ctor(**kw)
        sd = {k.split("/sd/", 1)[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix + "/sd/")}
        module.load_state_dict(sd, strict=True)
        _run_block(module, g, prefix)


def test_reference_inpainting_train_step_through_dropin():
    """BASELINE config 4: the reference's completion model (model_zoo/completion/inpainter.py, 53.4 M parameters, the
    AdaIN blocks at 16384 decoder points) and the training computation of train_inpainter.py:176-196 -- the reference's
    own partial_postproces, the EMD (0.005, 50) and the Chamfer loss -- with layers.cloud_transform,
    emd_linear.emd_module and chamfer_extension.dist_chamfer resolving to dropin/, every other file the reference's."""
    _need_tree()
    with RL.reference_tree(dropin=True) as rt:
        torch.manual_seed(0)
        generator = rt.load_model("model_zoo/completion/inpainter.py")
        import chamfer_extension.dist_chamfer as dist_chamfer
        import emd_linear.emd_module as emd
        from utils.pcd_utils import partial_postproces
        assert os.path.realpath(emd.__file__).startswith(os.path.realpath(RL.DROPIN_ROOT))
        assert os.path.realpath(dist_chamfer.__file__).startswith(os.path.realpath(RL.DROPIN_ROOT))
        assert abs(sum(p.numel() for p in generator.parameters()) / 1e6 - 53.41) < 0.01
        generator = generator.to(DEV).train()
        optimizer = torch.optim.Adam(generator.parameters(), lr=1e-4)
        EMD = emd.emdModule()
        B, n_in, n_gt = 2, 2048, 16384                           # configs/inpainting.yaml
        g = torch.Generator().manual_seed(3)
        u = torch.randn(B, n_gt, 3, generator=g)
        gt = 0.5 * u / u.norm(dim=-1, keepdim=True) * torch.tensor([1.0, 0.6, 0.4])      # points on an ellipsoid in [-0.5, 0.5]^3
        partial = gt[:, :n_in].clone()
        partial[:, 1500:] = 0.0                                  # zero rows = padding, dropped by partial_postproces
        losses = []
        for step in range(3):
            pcd_gt = 2 * gt.permute(0, 2, 1)[:, :, None].to(DEV)
            pcd_part_enc, pcd_part_noise = partial_postproces(2 * partial, pcd_gt.shape[-1])
            pcd_part_enc = pcd_part_enc.permute(0, 2, 1)[:, :, None].to(DEV)
            pcd_part_noise = pcd_part_noise.permute(0, 2, 1).to(DEV)
            reconstruction, lattices_sizes = generator(pcd_part_noise, pcd_part_enc)
            assert reconstruction.shape == (B, 3, 1, n_gt)
            dist, assignment = EMD(reconstruction[:, :, 0].permute(0, 2, 1), pcd_gt[:, :, 0].permute(0, 2, 1), 0.005, 50)
            loss_emd = torch.sqrt(dist).mean(1).mean()
            loss_chamfer = dist_chamfer.loss_chamfer(reconstruction, pcd_gt)
            loss = loss_emd + 0.1 * loss_chamfer
            loss.backward()
            grads = [p.grad for p in generator.parameters() if p.grad is not None]
            assert len(grads) > 100 and all(torch.isfinite(gr).all() for gr in grads)
            assert float(sum(gr.abs().sum() for gr in grads)) > 0.0
            optimizer.step()
            optimizer.zero_grad()
            assert int(assignment.min()) >= 0 and int(assignment.max()) < n_gt
            assert assignment.unique().numel() > n_gt // 2        # most of the targets are matched after 50 iterations
            losses.append(float(loss.detach()))
        assert all(np.isfinite(losses)), losses


def _load_pair(rel_path):
    """The same reference model file twice, same seed: on the reference's own torch ops and through dropin/."""
    with RL.reference_tree(dropin=False) as rt:
        torch.manual_seed(0)
        ref = rt.load_model(rel_path)
    with RL.reference_tree(dropin=True) as rt:
        torch.manual_seed(0)
        ours = rt.load_model(rel_path)
        import layers.cloud_transform as ct
        assert os.path.realpath(ct.__file__).startswith(os.path.realpath(RL.DROPIN_ROOT))
    ours.load_state_dict(ref.state_dict(), strict=True)
    return ref.to(DEV).train(), ours.to(DEV).train()


def _compare_models(ref, ours, make_args, pick, what):
    """Same input through both (train mode, dropout off, BatchNorm on batch statistics): outputs and gradients."""
    dropout_off(ref), dropout_off(ours)
    # (TF32 convolutions round their inputs to 10 bits: a 1e-7 difference between the two paths would come out as 1e-4)
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    outs = []
    for model in (ref, ours):
        args, leaf = make_args()
        out = pick(model(*args))
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(5)).to(DEV)
        (out * w).sum().backward()
        first = next(p for n, p in model.named_parameters() if n.startswith("first_process.0.weight"))
        outs.append((out.detach().cpu().numpy(), leaf.grad.cpu().numpy(), first.grad.cpu().numpy()))
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    close(outs[1][0], outs[0][0], what + ": output")
    close_l2(outs[1][1], outs[0][1], what + ": d loss / d input", rel=2e-2)
    close_l2(outs[1][2], outs[0][2], what + ": grad first conv", rel=2e-2)


def test_reference_s3dis_segmenter_through_dropin_matches_reference_ops_on_gpu():
    """BASELINE config 3: model_zoo/s3dis/segmenter.py (9.22 M parameters; train_segmentation.py:180-186, N = 4096, xyz +
    rgb) -- the reference's torch composition of Splat / Slice on the GPU against the B200 kernels, model level."""
    _need_tree()
    ref, ours = _load_pair("model_zoo/s3dis/segmenter.py")

    def make_args():
        g = torch.Generator().manual_seed(9)
        pcd = torch.cat([torch.rand(2, 2, 4096, generator=g), 3 * torch.rand(2, 1, 4096, generator=g),
                         torch.rand(2, 3, 4096, generator=g)], dim=1)[:, :, None].to(DEV).requires_grad_(True)
        return (pcd,), pcd

    _compare_models(ref, ours, make_args, lambda o: o[0], "s3dis segmenter")


def test_reference_padded_segmenter_through_dropin_matches_reference_ops_on_gpu():
    """model_zoo/s3dis/segmenter_pad.py: ragged clouds with a padding mask through every block (multihead_ct.py:84-107)."""
    _need_tree()
    ref, ours = _load_pair("model_zoo/s3dis/segmenter_pad.py")

    def make_args():
        g = torch.Generator().manual_seed(10)
        pts = (torch.rand(2, 3000, 3, generator=g) * 2 - 1).to(DEV)
        pad = torch.ones(2, 3000)
        pad[0, 2500:] = 0.0
        pad[1, 1777:] = 0.0
        feats = torch.rand(2, 4, 3000, generator=g).to(DEV).requires_grad_(True)
        return (pts, pad.to(DEV), feats), feats

    _compare_models(ref, ours, make_args, lambda o: o, "padded segmenter")


def test_reference_classifier_with_scales_through_dropin_matches_reference_ops_on_gpu():
    """model_zoo/scanobject/classifier_scales.py: the learnable per-axis scales of the lattice transforms."""
    _need_tree()
    ref, ours = _load_pair("model_zoo/scanobject/classifier_scales.py")

    def make_args():
        g = torch.Generator().manual_seed(11)
        pcd = (torch.rand(4, 3, 1, 2048, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
        return (pcd,), pcd

    _compare_models(ref, ours, make_args, lambda o: o[0], "classifier with scales")
