"""CPU: the PlaneTRHead oracle (oracle/planeTR_restate.py, row f1 first half) pinned against the LIVE reference module imported
from /root/reference (skipped where the reference tree is absent) and against the committed golden fixture generated from it;
plus the product's PlaneTRHead Python glue + plain-SIMT kernels (row op, tiled attention, top-down upsampling) executed on the
host (tests/simt_host) against the oracle — the tensor-engine GEMMs are functional stand-ins there (GPU tests hold the real ones)."""
import os

import pytest
import torch

from nopesac_b200 import synthetic
from oracle import planeTR_restate as R
from oracle import ref_planetr_loader as L
from tests import host_fixture, util

SHAPES = {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)}


def make_features(seed, N, H, W):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.relu(torch.randn(N, c, H // s, W // s, generator=g)) for k, (c, s) in SHAPES.items()}


def planetr_state(num_queries, seed=77):
    shapes = util.planetr_shapes(num_queries)
    return synthetic.make_weights(shapes, seed)


@pytest.mark.skipif(not L.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("NQ,N,H,W", [(20, 2, 96, 128), (50, 1, 160, 224)])
def test_restatement_matches_live_reference(NQ, N, H, W):
    head = L.build_head(NQ)
    assert {k: list(v.shape) for k, v in head.state_dict().items()} == {k: list(v) for k, v in util.planetr_shapes(NQ).items()}
    sd = planetr_state(NQ)
    head.load_state_dict(sd)
    feats = make_features(3, N, H, W)
    with torch.no_grad():
        want, hs = head(feats)
        got, hs2 = R.plane_tr_head(sd, feats)
    assert set(want) == set(got)
    # torch's nn.MultiheadAttention takes its fused eval fast path in the live module, the restatement calls
    # F.multi_head_attention_forward: same arithmetic, different association -> fp32 noise, not bit equality
    for k in want:
        assert util.maxdiff(want[k], got[k]) <= 1e-5 * max(1.0, float(want[k].abs().max())), k
    assert util.maxdiff(hs, hs2) <= 1e-5 * float(hs.abs().max())


def test_restatement_matches_golden():
    g = torch.load(os.path.join(util.GOLDEN_DIR, "planetr_nq20.golden"), weights_only=False)
    c = g["case"]
    sd = planetr_state(c["NQ"], c["weight_seed"])
    feats = make_features(c["feat_seed"], c["N"], c["H"], c["W"])
    with torch.no_grad():
        got, hs = R.plane_tr_head(sd, feats)
    for k, want in g["outputs"].items():
        v = hs if k == "query_feat" else got[k]
        assert util.maxdiff(v, want) <= 1e-5 * max(1.0, float(want.abs().max())), k


@pytest.fixture()
def host_ops(monkeypatch):
    return host_fixture.install(monkeypatch, with_tensor_standins=True)


def test_planetr_head_glue_on_host_matches_oracle(host_ops):
    from nopesac_b200 import config
    from nopesac_b200.meta_arch import RESNET50_OUTPUT_SHAPE
    from nopesac_b200.planeTR_head import build_planeTR_head, sine_position_table
    NQ, N, H, W = 20, 2, 96, 128
    cfg = config.inference_cfg(NQ, device="cpu")
    head = build_planeTR_head(cfg, RESNET50_OUTPUT_SHAPE)
    assert {k: list(v.shape) for k, v in head.state_dict().items()} == {k: list(v) for k, v in util.planetr_shapes(NQ).items()}
    sd = planetr_state(NQ)
    head.load_state_dict(sd)
    pos = sine_position_table(3, 4, 128)
    assert torch.equal(pos, R.position_embedding_sine(1, 3, 4, 128)[0].flatten(1).t().contiguous())
    feats = make_features(3, N, H, W)
    got, hs = head(feats)
    with torch.no_grad():
        want, hs2 = R.plane_tr_head(sd, feats)
    assert set(want) == set(got)
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert util.maxdiff(got[k], want[k]) <= 1e-4 * max(1.0, float(want[k].abs().max())), (k, util.maxdiff(got[k], want[k]))
    assert util.maxdiff(hs, hs2) <= 1e-4 * float(hs2.abs().max())
