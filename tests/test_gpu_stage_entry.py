"""GPU: whole stages behind ONE C-ABI call each (csrc/forward.cu; the entries SURVEY.md §8b lists for a host that does not want to
replicate the Python orchestration) — `nsac_match_forward` (MatchingHead: projection, 18 GNN layers, Sinkhorn, assignment) and
`nsac_refine_forward` (the one-plane RANSAC refinement K6 .. K10), `nsac_pixel_forward` (pixel pose network + AIM) and their
composition `nsac_head_forward` (= PlaneCameraHead.inference_Joint in one call).  The heads use them by default; NSAC_PY_STAGES=1 issues the
same launches from Python.  Both must give identical bits on every output, for every
selection rule, with an explicit hypothesis list, with an assignment override, and the entry must validate its arguments."""
import ctypes

import pytest
import torch

from tests import util
from tests.test_gpu_parity import _gpu

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]


def _flat(out):
    cams, tl, rl, lsp, ass, pro = out
    d = {f"cam.{k}.{kk}": v for k, c in cams.items() for kk, v in c.items()}
    d.update({f"ass.{k}": v for k, v in ass.items()})
    d.update({f"pro.{k}": v for k, v in pro.items() if torch.is_tensor(v)})
    d["lsp"] = lsp[0]
    return d


def _both(monkeypatch, NQ, cam, run):
    outs = {}
    for stage_entry in (True, False):
        if stage_entry:
            monkeypatch.delenv("NSAC_PY_STAGES", raising=False)
        else:
            monkeypatch.setenv("NSAC_PY_STAGES", "1")            # read by PlaneCameraHead.__init__
        head, match, sd, msd = util.build_cuda_heads(NQ, cam, 0.2, _gpu())
        assert head.use_stage_entry == stage_entry
        outs[stage_entry] = _flat(run(head, match))
        torch.cuda.synchronize()
    assert set(outs[True]) == set(outs[False])
    for k in outs[True]:
        assert torch.equal(outs[True][k], outs[False][k]), k
    return outs[True]


@pytest.mark.parametrize("cam", ["soft", "avg-all", "min-cost", "max-score"])
def test_stage_entry_equals_python_stages(monkeypatch, cam):
    dev = _gpu()
    from nopesac_b200 import synthetic
    NQ, P, B = 64, 8, 5
    b = synthetic.make_batch(7100, B, P).to(dev)
    poses = [util.initial_pose_for(7100 + i) for i in range(B)]
    ip = (torch.cat([p[0] for p in poses]).to(dev), torch.cat([p[1] for p in poses]).to(dev))
    out = _both(monkeypatch, NQ, cam, lambda head, match: head(None, None, b.planes1, b.planes2, b.app1, b.app2, matching_net=match,
                                                              initial_pose=ip))
    assert int(out["pro.matched_num"].min()) >= 1


def test_stage_entry_with_hypothesis_list_and_override(monkeypatch):
    dev = _gpu()
    from nopesac_b200 import synthetic
    NQ, P, B = 256, 16, 3
    b = synthetic.make_batch(7200, B, P).to(dev)
    poses = [util.initial_pose_for(7200 + i) for i in range(B)]
    ip = (torch.cat([p[0] for p in poses]).to(dev), torch.cat([p[1] for p in poses]).to(dev))
    hp = synthetic.all_pairs_hypotheses(P, NQ).to(dev, torch.int32)
    out = _both(monkeypatch, NQ, "soft", lambda head, match: head(None, None, b.planes1, b.planes2, b.app1, b.app2, matching_net=match,
                                                                 initial_pose=ip, hyp_pairs=hp))
    assert out["pro.matched_num"].cpu().tolist() == [NQ] * B
    ov = torch.zeros(B, P, P, device=dev)
    ov[:, torch.arange(P), torch.arange(P).flip(0)] = 1.0          # anti-diagonal: every plane "matched", mostly wrongly
    out = _both(monkeypatch, NQ, "soft", lambda head, match: head(None, None, b.planes1, b.planes2, b.app1, b.app2, matching_net=match,
                                                                 initial_pose=ip, assignment_override=ov))
    assert out["pro.matched_num"].cpu().tolist() == [P] * B


def test_single_call_head_with_pixel_network(monkeypatch):
    """From backbone feature maps (stage set S4): nsac_head_forward (K1 + K2 + matcher + refinement in one call) == the
    per-kernel Python path, bit for bit, and the result is within 1e-4 of the oracle for every pair."""
    dev = _gpu()
    from nopesac_b200 import synthetic
    from oracle import restate
    from tests.test_gpu_parity import _check_against
    NQ, P, B = 64, 8, 3
    b = synthetic.make_batch(7400, B, P, with_features=True)
    bd = b.to(dev)
    out = _both(monkeypatch, NQ, "soft", lambda head, match: head(bd.feats1, bd.feats2, bd.planes1, bd.planes2, bd.app1, bd.app2,
                                                                 matching_net=match))
    monkeypatch.delenv("NSAC_PY_STAGES", raising=False)
    head, match, sd, msd = util.build_cuda_heads(NQ, "soft", 0.2, dev)
    assert head.use_stage_entry and match.use_stage_entry
    cams, _, _, lsp, ass, pro = head(bd.feats1, bd.feats2, bd.planes1, bd.planes2, bd.app1, bd.app2, matching_net=match)
    torch.cuda.synchronize()
    assert torch.equal(pro["pose"], out["pro.pose"])
    with torch.no_grad():
        for i in range(B):
            o = restate.inference_joint(sd, msd, {k: v[i:i + 1] for k, v in b.feats1.items()}, {k: v[i:i + 1] for k, v in b.feats2.items()},
                                        b.planes1[i:i + 1], b.planes2[i:i + 1], b.app1[i:i + 1], b.app2[i:i + 1], num_queries=NQ)
            _check_against(util.oracle_to_flat(o), cams, lsp, ass, pro, i, f"single-call pair {i}")


def test_stage_entry_ragged_batch(monkeypatch):
    """Different plane counts per pair (count1 / count2 on the device) and n1 != n2: nsac_match_forward's ragged path."""
    dev = _gpu()
    from nopesac_b200 import synthetic
    NQ, B = 32, 4
    b = synthetic.make_batch(7300, B, 7).to(dev)
    poses = [util.initial_pose_for(7300 + i) for i in range(B)]
    ip = (torch.cat([p[0] for p in poses]).to(dev), torch.cat([p[1] for p in poses]).to(dev))
    c1 = torch.tensor([7, 3, 5, 1], dtype=torch.int32, device=dev)
    c2 = torch.tensor([6, 6, 2, 4], dtype=torch.int32, device=dev)
    p2, a2 = b.planes2[:, :6].contiguous(), b.app2[:, :6].contiguous()
    out = _both(monkeypatch, NQ, "soft", lambda head, match: head(None, None, b.planes1, p2, b.app1, a2, matching_net=match,
                                                                 initial_pose=ip, plane_count1=c1, plane_count2=c2))
    ass = out["ass.pred_assignment_beforeRef0"]
    for i in range(B):
        assert float(ass[i, int(c1[i]):].abs().sum()) == 0 and float(ass[i, :, int(c2[i]):].abs().sum()) == 0


@pytest.mark.parametrize("N,H,W", [(2, 97, 131), (2, 480, 640)])
def test_backbone_entry_equals_python_stages(monkeypatch, N, H, W):
    """nsac_backbone_forward (ResNet-50 from uint8 images in one call) == the per-kernel Python loop, every level, every bit."""
    dev = _gpu()
    from nopesac_b200 import backbone, config, synthetic
    images = synthetic.make_images(31, N, H, W).to(dev)
    outs = {}
    for stage_entry in (True, False):
        if stage_entry:
            monkeypatch.delenv("NSAC_PY_STAGES", raising=False)
        else:
            monkeypatch.setenv("NSAC_PY_STAGES", "1")
        net = backbone.build_backbone(config.inference_cfg())
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        net.load_state_dict(synthetic.make_backbone_weights(shapes, seed=8))
        net = net.to(dev)
        assert net.use_stage_entry == stage_entry
        outs[stage_entry] = net(images, planes=True)
        torch.cuda.synchronize()
    for k in ("res2", "res3", "res4", "res5"):
        (a, ha, wa), (b, hb, wb) = outs[True][k], outs[False][k]
        assert (ha, wa) == (hb, wb) and torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo), k


def test_model_entry_equals_python_stages(monkeypatch):
    """Stage set S5 behind ONE C call (nsac_model_forward: backbone on both views' uint8 images + the whole head) == backbone and
    head issued kernel by kernel from Python (NSAC_PY_STAGES=1), every output bit for bit; one launch counter for the whole step."""
    dev = _gpu()
    from nopesac_b200 import config, meta_arch, ops, synthetic
    NQ, P, B = 64, 8, 2
    b = synthetic.make_batch(7500, B, P).to(dev)
    images = synthetic.make_images(75, 2 * B).to(dev)
    outs, launches = {}, {}
    for stage_entry in (True, False):
        if stage_entry:
            monkeypatch.delenv("NSAC_PY_STAGES", raising=False)
        else:
            monkeypatch.setenv("NSAC_PY_STAGES", "1")
        model = meta_arch.PlaneTR_NopeSAC(config.inference_cfg(NQ), with_backbone=True)
        sd, msd = util.make_weights(NQ)
        model.camera_head_list[0].load_state_dict(sd)
        model.matching_head.load_state_dict(msd)
        shapes = {k: tuple(v.shape) for k, v in model.backbone.state_dict().items()}
        model.backbone.load_state_dict(synthetic.make_backbone_weights(shapes, seed=8))
        model = model.to(dev)
        assert model.backbone.use_stage_entry == stage_entry and model.camera_head_list[0].use_stage_entry == stage_entry
        n0 = ops.launch_count()
        outs[stage_entry] = _flat(model.inference_from_images(images, None, b.planes1, b.planes2, b.app1, b.app2))
        torch.cuda.synchronize()
        launches[stage_entry] = ops.launch_count() - n0
    assert set(outs[True]) == set(outs[False])
    for k in outs[True]:
        assert torch.equal(outs[True][k], outs[False][k]), k
    assert 300 < launches[True] <= launches[False] + 8        # (the C path splits the two views and builds the matcher pose with its own tiny kernels)


def test_stage_entry_argument_checks():
    dev = _gpu()
    from nopesac_b200 import _lib, ops
    head, match, *_ = util.build_cuda_heads(32, "soft", 0.2, dev)
    W = head.refine_weights()
    L = _lib.lib()
    assert L.nsac_refine_workspace_bytes(4, 32) > 0 and L.nsac_refine_workspace_bytes(0, 32) == 0
    z = lambda *s: torch.zeros(*s, device=dev)
    B, P, NQ = 2, 4, 32
    args = dict(planes1=z(B, P, 3) + 1, planes2=z(B, P, 3) + 1, assign=z(B, P, P), t0=z(B, 3), q0=z(B, 4), rot_feat0=z(B, 256),
                trans_feat0=z(B, 256))
    r = ops.refine_forward(W, num_queries=NQ, **args)                 # empty assignment: m = 0 -> the initial pose comes back
    torch.cuda.synchronize()
    assert r["matched_num"].cpu().tolist() == [0, 0] and torch.isfinite(r["pose"]).all()
    # too small a workspace / a null weight struct are refused with a message, nothing is launched
    st = L.nsac_refine_forward(ctypes.byref(W), *[ctypes.c_void_p(args[k].data_ptr()) for k in ("planes1", "planes2", "assign")], None, 0,
                               *[ctypes.c_void_p(args[k].data_ptr()) for k in ("t0", "q0", "rot_feat0", "trans_feat0")], B, P, P, NQ, 0,
                               *[ctypes.c_void_p(r[k].data_ptr()) if r[k] is not None else None for k in
                                 ("pose", "assign_pruned", "geo_local", "geo_global", "sig", "matched_num", "pair_idx", "q_h", "t_h",
                                  "score_rot", "score_tran", "sel_idx")],
                               ctypes.c_void_p(r["pose"].data_ptr()), 64, None, 0, 0, None, None)
    assert st != 0 and b"workspace" in L.nsac_last_error()
