"""Host-side mirror of the reference model API: state-dict contract, builders, flag rejection (CPU only)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lavt_rs_b200.args import default_args  # noqa: E402
from oracle import lavt_oracle as O  # noqa: E402
from oracle import ref_shims  # noqa: E402


def _small_backbone(window=(8, 7, 7)):
    from lavt_rs_b200.lib.mask_predictor import SimpleDecoding
    from lavt_rs_b200.lib.video_swin_transformer import MultiModalSwinTransformer3D
    bb = MultiModalSwinTransformer3D(patch_size=(1, 4, 4), embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32],
                                     window_size=window, drop_path_rate=0.0, patch_norm=True, args=None)
    return bb, SimpleDecoding(1024, None)


def test_state_dict_keys_match_reference_contract():
    bb, dec = _small_backbone()
    cfg = O.OracleConfig(depths=(2, 2, 2, 2))
    sd = O.random_state_dict(cfg)
    mine = {"backbone." + k: v for k, v in bb.state_dict().items()}
    mine.update({"classifier." + k: v for k, v in dec.state_dict().items()})
    benign = ("relative_position_index", "num_batches_tracked")
    assert {k for k in mine if not k.endswith(benign)} == set(sd)
    for k, v in sd.items():
        assert tuple(mine[k].shape) == tuple(v.shape), k
    # the buffers the reference registers are present too (strict loading of reference checkpoints)
    assert mine["backbone.layers.0.blocks.0.attn.relative_position_index"].shape == (392, 392)
    assert "classifier.bn1_4.num_batches_tracked" in mine


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present")
def test_full_model_state_dict_round_trips_with_reference():
    from lavt_rs_b200.lib import segmentation
    ref, _ = ref_shims.build_reference("lavt_video", "base")
    mine = segmentation.lavt_video(pretrained="", args=default_args(["--model", "lavt_video", "--swin_type", "base"]))
    ref_sd = ref.state_dict()
    my_sd = mine.state_dict()
    strip = lambda d: {k for k in d if not k.startswith("text_encoder.")}
    assert strip(ref_sd) == strip(my_sd)
    for k in strip(ref_sd):
        assert ref_sd[k].shape == my_sd[k].shape, k
    mine.load_state_dict({k: v for k, v in ref_sd.items() if not k.startswith("text_encoder.")}, strict=False)
    idx = "backbone.layers.2.blocks.5.attn.relative_position_index"
    assert torch.equal(mine.state_dict()[idx], ref_sd[idx])
    # method wiring of the 2-D -> video checkpoint inflation (the tensor math is pinned against the reference's own
    # loader in tests/test_oracle_vs_reference.py): a synthetic 2-D checkpoint cut out of the video state dict
    import contextlib
    import io
    import tempfile
    sd2d = {k: v.clone() for k, v in ref_sd.items() if not k.startswith("text_encoder.")}
    sd2d["backbone.patch_embed.proj.weight"] = sd2d["backbone.patch_embed.proj.weight"].squeeze(2)
    tkey = "backbone.layers.1.blocks.0.attn.relative_position_bias_table"
    for k in [k for k in sd2d if "relative_position_bias_table" in k]:
        sd2d[k] = sd2d[k][:169] + 1.0
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "lavt2d.pth")
        torch.save({"model": sd2d}, path)
        with contextlib.redirect_stdout(io.StringIO()):
            mine.load_from_pretrained2d_lavt_weights(path)
    got = mine.state_dict()
    assert got["backbone.patch_embed.proj.weight"].shape == ref_sd["backbone.patch_embed.proj.weight"].shape
    assert torch.equal(got[tkey], sd2d[tkey].repeat(15, 1))


def test_builders_reject_unimplemented_flags():
    from lavt_rs_b200.lib import segmentation
    for flag in ("--sep_t_pwam", "--ts_pwam"):       # (--sep_t_pwam alone keeps the unsupported 3-1-1 temporal kernel)
        with pytest.raises(NotImplementedError):
            segmentation.lavt_video(pretrained="", args=default_args(["--swin_type", "base", flag]))
    # Swin-T / Swin-S widths (96 channels) build: the README's video commands use --swin_type tiny
    tiny = segmentation.lavt_video(pretrained="", args=default_args(["--swin_type", "tiny"]))
    assert tiny.backbone.embed_dim == 96


def test_bcam_gacd_backbone_state_dicts_match_oracle_contract():
    """--bcam / --efn / --gacd (reference lib/backbone.py:573-588): the 2-D backbone registers the reference's fusion parameters."""
    from lavt_rs_b200.lib.backbone import MultiModalSwinTransformer
    for flag in ("bcam", "efn", "gacd"):
        bb = MultiModalSwinTransformer(embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32], window_size=7, drop_path_rate=0.0,
                                       num_heads_fusion=[1, 1, 1, 1], args=default_args(["--" + flag]))
        cfg = O.OracleConfig(depths=(2, 2, 2, 2), window=(1, 7, 7), clamp_window=False, video=False, **{flag: True})
        sd = {k[len("backbone."):]: v for k, v in O.random_state_dict(cfg).items() if k.startswith("backbone.")}
        mine = {k: v for k, v in bb.state_dict().items() if not k.endswith("relative_position_index")}
        assert set(mine) == set(sd), flag
        for k, v in sd.items():
            assert tuple(mine[k].shape) == tuple(v.shape), k
    assert bb.layers[0].fusion.kind == "gacd"


def test_lazy_pred_state_dict_matches_oracle_contract():
    """--lazy_pred (reference lib/segmentation.py:183-185, lib/mask_predictor.py:32): no backbone.norm0, no 1/4-scale decoder level."""
    from lavt_rs_b200.lib import segmentation
    m = segmentation.lavt_video(pretrained="", args=default_args(["--swin_type", "tiny", "--lazy_pred"]))
    assert m.backbone.out_indices == (1, 2, 3) and m.lazy_pred and m.classifier.lazy_pred and m.backbone.layers[0].lazy_pred
    cfg = O.OracleConfig(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), lazy_pred=True)
    sd = O.random_state_dict(cfg)
    mine = {k for k in m.state_dict() if not k.startswith("text_encoder.") and not k.endswith(("relative_position_index", "num_batches_tracked"))}
    assert mine == set(sd)
    assert "backbone.norm0.weight" not in mine and "classifier.conv1_2.weight" not in mine


def test_inference_only_variants_refuse_training():
    """The lib/bcam.py fusions have no hand-written backward: the training entry point must refuse them up front.  --lazy_pred,
    --fuse simple, the BN / LN / none attention norms and the decoder tails train (tests/test_backward_gpu.py)."""
    from lavt_rs_b200 import training
    from lavt_rs_b200.lib import segmentation
    for flag in (["--bcam"], ["--efn"], ["--gacd"]):
        m = segmentation.lavt(pretrained="", args=default_args(["--swin_type", "base", *flag]))
        with pytest.raises(NotImplementedError):
            training._check_trainable(m)
    training._check_trainable(segmentation.lavt(pretrained="", args=default_args(["--swin_type", "tiny"])))
    training._check_trainable(segmentation.lavt(pretrained="", args=default_args(["--swin_type", "base", "--lazy_pred"])))
    training._check_trainable(segmentation.lavt(pretrained="", args=default_args(["--swin_type", "base", "--interpolate_before_seg", "--seg_last"])))
    for kind in ("none", "LN", "BN"):
        training._check_trainable(segmentation.lavt(pretrained="", args=default_args(["--swin_type", "base", "--att_norm_layer_type", kind])))
    training._check_trainable(segmentation.lavt_video(pretrained="", args=default_args(["--swin_type", "tiny", "--fuse", "simple"])))
    # the sigmoid gate trains (gate adjoint modes 7 / 8 of lavt_gate_elementwise)
    training._check_trainable(segmentation.lavt(pretrained="", args=default_args(["--swin_type", "base", "--lg_act_layer", "sigmoid"])))


def test_forward_refuses_cpu_tensors():
    from lavt_rs_b200 import _cabi
    bb, dec = _small_backbone()
    with pytest.raises(_cabi.LavtError):
        bb(torch.zeros(1, 3, 2, 32, 32), torch.zeros(1, 768, 4), torch.ones(1, 4, 1))
    with pytest.raises(_cabi.LavtError):
        dec(torch.zeros(1, 1024, 1, 1), torch.zeros(1, 512, 2, 2), torch.zeros(1, 256, 4, 4), torch.zeros(1, 128, 8, 8))


def test_relative_position_index_closed_form():
    from lavt_rs_b200.lib.video_swin_transformer import _relative_position_index
    for window in ((8, 7, 7), (2, 3, 4), (1, 12, 12)):
        Wd, Wh, Ww = window
        co = torch.stack(torch.meshgrid(torch.arange(Wd), torch.arange(Wh), torch.arange(Ww), indexing="ij")).flatten(1)
        rel = (co[:, :, None] - co[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += Wd - 1
        rel[:, :, 1] += Wh - 1
        rel[:, :, 2] += Ww - 1
        rel[:, :, 0] *= (2 * Wh - 1) * (2 * Ww - 1)
        rel[:, :, 1] *= 2 * Ww - 1
        assert torch.equal(_relative_position_index(window), rel.sum(-1))


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("model", ["vlt", "lavt_vlt"])
def test_vlt_builders_state_dict_matches_reference(model):
    """``segmentation.vlt`` / ``segmentation.lavt_vlt`` (reference lib/segmentation.py:299-433): same classes of sub-modules, same
    state-dict keys and shapes as the reference's own builders, so that its checkpoints load unchanged."""
    from lavt_rs_b200.lib import segmentation
    extra = ["--img_size", "480"]
    ref, _ = ref_shims.build_reference(model, "base", extra=extra)
    mine = segmentation.__dict__[model](pretrained="", args=default_args(["--model", model, "--swin_type", "base", *extra]))
    ref_sd, my_sd = ref.state_dict(), mine.state_dict()
    strip = lambda d: {k for k in d if not k.startswith("text_encoder.")}   # noqa: E731
    assert strip(ref_sd) == strip(my_sd), (sorted(strip(ref_sd) - strip(my_sd))[:5], sorted(strip(my_sd) - strip(ref_sd))[:5])
    for k in strip(ref_sd):
        assert ref_sd[k].shape == my_sd[k].shape, k
    assert type(mine).__name__ == type(ref).__name__
    missing = mine.load_state_dict({k: v for k, v in ref_sd.items() if not k.startswith("text_encoder.")}, strict=False)
    assert not missing.unexpected_keys and all(k.startswith("text_encoder.") for k in missing.missing_keys)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("extra", [["--att_norm_layer_type", "BN"], ["--att_norm_layer_type", "LN"], ["--att_norm_layer_type", "none"],
                                   ["--lg_act_layer", "sigmoid"], ["--interpolate_before_seg"], ["--interpolate_before_seg", "--seg_last"]])
def test_flag_variants_state_dict_matches_reference(extra):
    """2-D gate / norm options and the extra decoder levels: same state-dict keys and shapes as the reference builder with the same flags."""
    from lavt_rs_b200.lib import segmentation
    ref, _ = ref_shims.build_reference("lavt_one", "tiny", extra=extra)
    mine = segmentation.lavt_one(pretrained="", args=default_args(["--model", "lavt_one", "--swin_type", "tiny", *extra]))
    strip = lambda d: {k: v.shape for k, v in d.items() if not k.startswith("text_encoder.")}   # noqa: E731
    assert strip(ref.state_dict()) == strip(mine.state_dict())
    if extra[0] == "--lg_act_layer":
        assert type(mine.backbone.layers[0].res_gate[3]).__name__ == type(ref.backbone.layers[0].res_gate[3]).__name__ == "Sigmoid"


def test_image_backbone_initialises_from_swin_checkpoint(tmp_path):
    """``init_weights(pretrained=path)`` of the 2-D backbone (reference lib/backbone.py:476-486: OpenMMLab load_checkpoint, strict=False):
    the Swin keys of the file are loaded ('module.' prefix stripped, 'state_dict' / 'model' wrappers accepted), fusion / gate / per-stage
    norm parameters keep their initialisation."""
    from lavt_rs_b200.lib.backbone import MultiModalSwinTransformer
    kw = dict(embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32], window_size=7, drop_path_rate=0.0, patch_norm=True,
              num_heads_fusion=[1, 1, 1, 1], args=None)
    src = MultiModalSwinTransformer(**kw)
    src.init_weights()
    swin = {"module." + k: v + 0.5 for k, v in src.state_dict().items()
            if v.is_floating_point() and "fusion" not in k and "res_gate" not in k and not k.startswith("norm")}
    path = str(tmp_path / "swin_imagenet.pth")
    torch.save({"model": swin}, path)
    dst = MultiModalSwinTransformer(**kw)
    torch.manual_seed(1)
    dst.init_weights(pretrained=path)
    got = dst.state_dict()
    for k, v in swin.items():
        assert torch.equal(got[k[len("module."):]], v), k
    assert got["layers.0.fusion.vis_project.0.weight"].abs().sum() > 0 and torch.equal(got["norm0.weight"], torch.ones(128))
