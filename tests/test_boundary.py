"""CPU: the drop-in boundary — registries, config loading, state-dict names."""
import os

import pytest
import torch

from nopesac_b200 import (CAMERA_HEAD_REGISTRY, MATCHING_HEAD_REGISTRY, META_ARCH_REGISTRY, build_camera_head,
                          build_matching_head, config)
from nopesac_b200.compat import CfgNode, Registry
from nopesac_b200.meta_arch import RESNET50_OUTPUT_SHAPE, PlaneTR_NopeSAC
from tests import util

REF_CFG = "/root/reference/configs/inference_mp3d.yaml"


def test_registries_resolve_reference_names():
    assert CAMERA_HEAD_REGISTRY.get("PlaneCameraHead").__name__ == "PlaneCameraHead"
    assert MATCHING_HEAD_REGISTRY.get("MatchingHead").__name__ == "MatchingHead"
    assert META_ARCH_REGISTRY.get("PlaneTR_NopeSAC") is PlaneTR_NopeSAC
    with pytest.raises(KeyError):
        CAMERA_HEAD_REGISTRY.get("Nope")
    r = Registry("X")

    @r.register()
    class A:  # noqa: D401
        pass
    with pytest.raises(AssertionError):
        r.register(A)


def test_cfgnode_base_inheritance_and_overrides(tmp_path):
    (tmp_path / "Base.yaml").write_text("MODEL:\n  RESNETS:\n    DEPTH: 50\nSOLVER:\n  OPTIMIZER: \"ADAMW\"\nVERSION: 2\n")
    (tmp_path / "child.yaml").write_text(
        "_BASE_: Base.yaml\nMODEL:\n  CAMERA_HEAD:\n    NAME: \"PlaneCameraHead\"\n    REFINE_ON: True\n"
        "DATASETS:\n  TEST: (\"mp3d_test\",)\nTEST:\n  MATCHING_SCORE_THRESHOLD: 0.2\n")
    cfg = config.load_config(str(tmp_path / "child.yaml"),
                             ["MODEL.SEM_SEG_HEAD.NUM_OBJECT_QUERIES", "256", "MODEL.DEVICE", "cpu"])
    assert cfg.MODEL.CAMERA_HEAD.NAME == "PlaneCameraHead" and cfg.MODEL.CAMERA_HEAD.REFINE_ON is True
    assert cfg.MODEL.RESNETS.DEPTH == 50 and cfg.DATASETS.TEST == ("mp3d_test",)
    assert cfg.MODEL.SEM_SEG_HEAD.NUM_OBJECT_QUERIES == 256 and cfg.MODEL.DEVICE == "cpu"
    assert cfg.MODEL.MATCHING_HEAD.OFFSET_MULTIPLIER == 4.0       # NopeSAC default survives
    with pytest.raises(AttributeError):
        cfg.MODEL.DEVICE = "cuda"                                  # frozen
    with pytest.raises(KeyError):
        c2 = cfg.clone(); c2.defrost(); c2.merge_from_list(["MODEL.NOPE", 1])
    with pytest.raises(ValueError):
        c3 = cfg.clone(); c3.defrost(); c3.merge_from_list(["MODEL.CAMERA_HEAD.NAME", 3])


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not mounted")
def test_reference_yaml_loads_unchanged():
    cfg = config.load_config(REF_CFG)
    assert cfg.MODEL.META_ARCHITECTURE == "PlaneTR_NopeSAC"
    assert cfg.MODEL.CAMERA_HEAD.NAME == "PlaneCameraHead" and cfg.MODEL.CAMERA_HEAD.INFERENCE_OUT_CAM_TYPE == "soft"
    assert cfg.MODEL.CAMERA_HEAD.CAM_REC_ON and cfg.MODEL.CAMERA_HEAD.REFINE_ON and cfg.MODEL.EMBEDDING_ON
    assert cfg.MODEL.SEM_SEG_HEAD.NORM == "GN" and cfg.MODEL.SEM_SEG_HEAD.NUM_OBJECT_QUERIES == 50
    assert cfg.TEST.MATCHING_SCORE_THRESHOLD == 0.2 and cfg.SOLVER.BACKBONE_MULTIPLIER == 0.1
    same = config.inference_cfg(50)
    for path in ("MODEL.CAMERA_HEAD", "MODEL.MATCHING_HEAD", "TEST"):
        a, b = cfg, same
        for p in path.split("."):
            a, b = a[p], b[p]
        assert dict(a) == dict(b), path
    head = build_camera_head(cfg, RESNET50_OUTPUT_SHAPE)
    assert head.num_queries == 50 and head.inference_out_cam_type == "soft"


@pytest.mark.parametrize("nq", [50, 256])
def test_state_dict_names_and_shapes_match_reference(nq):
    """Names/shapes recorded from the live reference modules by tests/golden/make_golden.py."""
    cfg = config.inference_cfg(nq, device="cpu")
    head = build_camera_head(cfg, RESNET50_OUTPUT_SHAPE)
    match = build_matching_head(cfg)
    hs, ms = util.state_shapes(nq)
    assert {k: list(v.shape) for k, v in head.state_dict().items()} == {k: list(v) for k, v in hs.items()}
    assert {k: list(v.shape) for k, v in match.state_dict().items()} == {k: list(v) for k, v in ms.items()}
    assert list(head.state_dict()) == list(hs) or sorted(head.state_dict()) == sorted(hs)


def test_meta_arch_loads_reference_checkpoint_keys():
    cfg = config.inference_cfg(50, device="cpu")
    model = PlaneTR_NopeSAC(cfg)
    hs, ms = util.make_weights(50)
    ckpt = {"camera_head_list.0." + k: v for k, v in hs.items()}
    ckpt.update({"matching_head." + k: v for k, v in ms.items()})
    ckpt["backbone.stem.conv1.weight"] = torch.zeros(64, 3, 7, 7)       # a key of the out-of-scope detector
    ignored = model.load_reference_state_dict(ckpt)
    assert ignored == ["backbone.stem.conv1.weight"]
    assert torch.equal(model.camera_head_list[0].rots.weight, hs["rots.weight"])
    assert torch.equal(model.matching_head.gnn.layers[17].mlp[2].weight, ms["gnn.layers.17.mlp.2.weight"])


def test_weight_packs_are_invalidated_on_load():
    cfg = config.inference_cfg(50, device="cpu")
    head = build_camera_head(cfg, RESNET50_OUTPUT_SHAPE)
    pk = head.prepare()
    assert pk["decoder_rot2.w_init"].shape == (512, 256)
    hs, _ = util.make_weights(50)
    head.load_state_dict(hs)
    assert head._packed is None
    assert torch.equal(head.prepare()["decoder_rot2.w_geo"], hs["decoder_rot2.layers.0.weight"][:, 256:])


def test_stack_views_pads_ragged_pairs_and_reports_counts():
    """Host logic of PlaneTR_NopeSAC.inference for pairs with different plane counts: zero padding to the largest count,
    int32 counts (None when the view is not ragged), feature maps stacked."""
    import torch
    g = torch.Generator().manual_seed(0)
    mk = lambda n: {"pred_plane": torch.randn(n, 3, generator=g), "pred_plane_feats": torch.randn(1, n, 256, generator=g),
                    "cam_feats": {"res5": torch.randn(1, 8, 2, 3, generator=g)}}
    bis = [{"0": mk(4), "1": mk(5)}, {"0": mk(2), "1": mk(5)}, {"0": mk(7), "1": mk(5)}]
    p, f, cam, cnt = PlaneTR_NopeSAC._stack_views(bis, "0", torch.device("cpu"))
    assert p.shape == (3, 7, 3) and f.shape == (3, 7, 256) and cnt.dtype == torch.int32 and cnt.tolist() == [4, 2, 7]
    assert torch.equal(p[1, :2], bis[1]["0"]["pred_plane"]) and float(p[1, 2:].abs().sum()) == 0.0
    assert torch.equal(f[0, :4], bis[0]["0"]["pred_plane_feats"][0]) and float(f[0, 4:].abs().sum()) == 0.0
    assert cam["res5"].shape == (3, 8, 2, 3)
    p, f, cam, cnt = PlaneTR_NopeSAC._stack_views(bis, "1", torch.device("cpu"))
    assert p.shape == (3, 5, 3) and cnt is None
    import pytest
    bad = [{"0": {"pred_plane": torch.zeros(0, 3), "pred_plane_feats": torch.zeros(0, 256), "cam_feats": {}}}]
    with pytest.raises(ValueError):
        PlaneTR_NopeSAC._stack_views(bad, "0", torch.device("cpu"))
