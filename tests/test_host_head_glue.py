"""CPU: the product's Python glue — `PlaneCameraHead.inference_Joint` (pixel pose CNN included) + `MatchingHead.match` —
executed end to end on CPU tensors against the golden fixtures generated from the live reference and the oracle.
The plain-SIMT kernels run from their real source on the host (tests/simt_host); the tensor-engine entry points are functional
stand-ins (tests/simt_host/tc_standin.cpp: hi/lo-plane GEMM in double, scoring routed to the exact-fp32 kernel).  What this
covers in a container without a GPU: weight packing, plane formats, row ranges / column slices of the GNN token buffer, geo
sequences, hypothesis features, per-pair m rules, result layout — the same `_check_against` bar as tests/test_gpu_parity.py.
What it does NOT cover: the tcgen05 kernels themselves (GPU tests)."""
import pytest
import torch

from nopesac_b200 import ops
from tests import host_fixture, util
from tests.test_gpu_parity import _check_against, _selection_from_reference

# The host execution is slow (every warp shuffle is two 32-thread barriers): by default three fixtures, first pair of each —
# c1_nq32_p8 = BASELINE.json configs[0], the reference's own CPU case, WITH backbone feature maps (the pixel pose CNN K1 runs on
# the host too: its glue kernels from source, the convolutions through the stand-in, i.e. the whole `inference_Joint`), ragged
# plane counts n1 != n2, and an index-selection mode.  NSAC_HOST_GLUE_ALL=1 runs every fixture with NQ <= 64, all pairs
# (about ten minutes).
import os

_ALL = [n for n in util.golden_names() if util.load_golden(n)["case"]["NQ"] <= 64]
_FAST = ["c1_nq32_p8", "ragged_nq50_p5x9", "maxscore_nq50_p16"]
NO_FEATS = _ALL if os.environ.get("NSAC_HOST_GLUE_ALL") == "1" else [n for n in _FAST if n in _ALL]
MAX_PAIRS = None if os.environ.get("NSAC_HOST_GLUE_ALL") == "1" else 1


@pytest.fixture()
def host_ops(monkeypatch):
    return host_fixture.install(monkeypatch, with_tensor_standins=True)


def _run_host_case(case, pair_indices, want_diag=False, **kw):
    head, match, sd, msd = util.build_cuda_heads(case["NQ"], case["cam"], case["thr"], "cpu")
    bs = [util.case_batch(case, pi) for pi in pair_indices]
    cat = lambda f: torch.cat([f(b) for b in bs], 0)
    p1, p2, a1, a2 = cat(lambda b: b.planes1), cat(lambda b: b.planes2), cat(lambda b: b.app1), cat(lambda b: b.app2)
    f1 = f2 = ip = None
    if case["feats"]:
        f1 = {k: torch.cat([b.feats1[k] for b in bs], 0) for k in bs[0].feats1}
        f2 = {k: torch.cat([b.feats2[k] for b in bs], 0) for k in bs[0].feats2}
    else:
        poses = [util.initial_pose_for(pi) for pi in pair_indices]
        ip = (torch.cat([p[0] for p in poses]), torch.cat([p[1] for p in poses]))
    hp = util.case_hyp_pairs(case)
    hp = None if hp is None else hp.to(torch.int32)
    return head(f1, f2, p1, p2, a1, a2, matching_net=match, hyp_pairs=hp, initial_pose=ip, want_diag=want_diag, **kw)


@pytest.mark.parametrize("name", NO_FEATS)
def test_head_glue_on_host_matches_golden(host_ops, name):
    g = util.load_golden(name)
    case, pairs = g["case"], g["case"]["pairs"][:MAX_PAIRS]
    cams, _, _, lsp, ass, pro = _run_host_case(case, pairs, want_diag=True)
    for i, (pi, want) in enumerate(zip(pairs, g["outputs"][:MAX_PAIRS])):
        _check_against(want, cams, lsp, ass, pro, i, f"{name}[pair {pi}] (host)")
        if case["cam"] in ("min-cost", "max-score") and int(want["matched_num"]) > 1:
            assert pro["sel_idx"][i].tolist() == list(_selection_from_reference(want, case)), name


def _flat(out):
    cams, tl, rl, lsp, ass, pro = out
    d = {f"cam.{k}.{kk}": v for k, c in cams.items() for kk, v in c.items()}
    d.update({f"ass.{k}": v for k, v in ass.items()})
    d.update({f"pro.{k}": v for k, v in pro.items() if torch.is_tensor(v)})
    d["lsp"] = lsp[0]
    return d


@pytest.mark.parametrize("name", NO_FEATS[:2])
def test_stage_entry_equals_python_stages_on_host(host_ops, name, monkeypatch):
    """K6 .. K10 through the single C entry (nsac_refine_forward, csrc/forward.cu compiled for the host) == the same launches
    issued from Python (head.use_stage_entry = False): every output bit-identical, incl. an assignment override."""
    g = util.load_golden(name)
    case, pairs = g["case"], g["case"]["pairs"][:MAX_PAIRS]
    outs = {}
    for flag in (True, False):
        if flag:
            monkeypatch.delenv("NSAC_PY_STAGES", raising=False)
        else:
            monkeypatch.setenv("NSAC_PY_STAGES", "1")          # read by PlaneCameraHead.__init__
        n0 = ops.launch_count()
        outs[flag] = _flat(_run_host_case(case, pairs))
        assert ops.launch_count() - n0 > 300          # the stage entries report the kernels they enqueue
    assert set(outs[True]) == set(outs[False])
    for k in outs[True]:
        assert torch.equal(outs[True][k], outs[False][k]), (name, k)
    for i, (pi, want) in enumerate(zip(pairs, g["outputs"][:MAX_PAIRS])):
        assert int(outs[True]["pro.matched_num"][i]) == int(want["matched_num"]), (name, pi)


def test_ragged_batch_glue_on_host_matches_per_pair_oracle(host_ops):
    """Pairs with DIFFERENT plane counts in one padded batch (`plane_count1/2`, the hand-off format of row f1's PlaneLists):
    every pair equals the oracle run on its un-padded slice; the padding (NaN here) is never read; outputs outside a pair's
    block are 0 / -inf."""
    from oracle import restate
    NQ, P = 32, 6
    counts = [(6, 6), (3, 5), (4, 2)]
    B = len(counts)
    head, match, sd, msd = util.build_cuda_heads(NQ, "soft", 0.2, "cpu")
    from nopesac_b200 import synthetic
    b = synthetic.make_batch(60, B, P)
    poses = [util.initial_pose_for(200 + i) for i in range(B)]
    ip = (torch.cat([p[0] for p in poses]), torch.cat([p[1] for p in poses]))
    p1, p2, a1, a2 = b.planes1.clone(), b.planes2.clone(), b.app1.clone(), b.app2.clone()
    outs = []
    with torch.no_grad():
        for i, (n1, n2) in enumerate(counts):
            outs.append(restate.inference_joint(sd, msd, None, None, p1[i:i + 1, :n1], p2[i:i + 1, :n2], a1[i:i + 1, :n1], a2[i:i + 1, :n2],
                                                num_queries=NQ, initial_pose=(ip[0][i:i + 1], ip[1][i:i + 1])))
            p1[i, n1:], a1[i, n1:], p2[i, n2:], a2[i, n2:] = float("nan"), float("nan"), float("nan"), float("nan")
    c1 = torch.tensor([c[0] for c in counts], dtype=torch.int32)
    c2 = torch.tensor([c[1] for c in counts], dtype=torch.int32)
    cams, _, _, lsp, ass, pro = head(None, None, p1, p2, a1, a2, matching_net=match, initial_pose=ip, plane_count1=c1, plane_count2=c2)
    assert sum(o["matched_num"] for o in outs) > 0
    for i, ((n1, n2), o) in enumerate(zip(counts, outs)):
        cams_i = {k: {"tran": v["tran"][i:i + 1], "rot": v["rot"][i:i + 1]} for k, v in cams.items() if v["tran"].shape[0] == B}
        ass_i = {k: v[i:i + 1, :n1, :n2] for k, v in ass.items()}
        pro_i = {k: v[i:i + 1] for k, v in pro.items() if torch.is_tensor(v) and v.shape[0] == B}
        _check_against(util.oracle_to_flat(o), cams_i, [lsp[0][i:i + 1, :n1 + 1, :n2 + 1]], ass_i, pro_i, 0, f"ragged pair {i} ({n1}x{n2})")
        for k, v in ass.items():
            assert float(v[i, n1:].abs().sum()) == 0.0 and float(v[i, :, n2:].abs().sum()) == 0.0, (i, k)
        assert bool(torch.isinf(lsp[0][i, n1 + 1:]).all()) and bool(torch.isinf(lsp[0][i, :, n2 + 1:]).all())
    for v in cams.values():
        assert bool(torch.isfinite(v["tran"]).all()) and bool(torch.isfinite(v["rot"]).all())


def test_plane_lists_to_camera_head_glue_on_host(host_ops):
    """`PlaneTR_NopeSAC.inference_from_plane_heads` (row f1 plane lists -> camera head with per-pair plane counts, no host
    round trip) on CPU tensors: plane lists equal the post-processing oracle's, poses / assignments equal the head oracle run
    pair by pair on those un-padded lists."""
    from nopesac_b200 import config, meta_arch, synthetic
    from oracle import planes_restate, restate
    NQ, h, w = 12, 9, 13
    H, W = 4 * h, 4 * w
    model = meta_arch.PlaneTR_NopeSAC(config.inference_cfg(NQ, device="cpu"))
    sd, msd = util.make_weights(NQ)
    model.camera_head_list[0].load_state_dict(sd)
    model.matching_head.load_state_dict(msd)
    B = 2
    kw = dict(num_queries=NQ, mask_h=h, mask_w=w, channels=256)
    v1 = synthetic.make_plane_head_batch(700, B, cases=("regular",), planes=5, **kw)
    v2 = synthetic.make_plane_head_batch(800, B, cases=("regular", "zero"), planes=4, **kw)
    poses = [util.initial_pose_for(300 + i) for i in range(B)]
    ip = (torch.cat([p[0] for p in poses]), torch.cat([p[1] for p in poses]))
    outs = lambda v: {k: v[k] for k in ("pred_logits", "pred_params", "pred_mask_logits")}
    (cams, _, _, lsp, ass, pro), l1, l2 = model.inference_from_plane_heads(outs(v1), v1["query_feat"], outs(v2), v2["query_feat"],
                                                                           None, None, height=H, width=W, max_planes=8, initial_pose=ip)
    o1 = planes_restate.postprocess_plane_head_mask(v1["pred_logits"], v1["pred_params"], v1["pred_mask_logits"], v1["query_feat"], H, W)
    o2 = planes_restate.postprocess_plane_head_mask(v2["pred_logits"], v2["pred_params"], v2["pred_mask_logits"], v2["query_feat"], H, W)
    assert l1.count.tolist() == [len(o["pred_plane_oriIdxs"]) for o in o1] and l2.count.tolist() == [len(o["pred_plane_oriIdxs"]) for o in o2]
    assert max(l1.count.max(), l2.count.max()) <= 8 and l1.count.tolist() != l2.count.tolist()
    for i in range(B):
        n1, n2 = int(l1.count[i]), int(l2.count[i])
        with torch.no_grad():
            o = restate.inference_joint(sd, msd, None, None, o1[i]["pred_plane"][None], o2[i]["pred_plane"][None], o1[i]["pred_plane_feats"],
                                        o2[i]["pred_plane_feats"], num_queries=NQ, initial_pose=(ip[0][i:i + 1], ip[1][i:i + 1]))
        cams_i = {k: {"tran": v["tran"][i:i + 1], "rot": v["rot"][i:i + 1]} for k, v in cams.items() if v["tran"].shape[0] == B}
        ass_i = {k: v[i:i + 1, :n1, :n2] for k, v in ass.items()}
        pro_i = {k: v[i:i + 1] for k, v in pro.items() if torch.is_tensor(v) and v.shape[0] == B}
        _check_against(util.oracle_to_flat(o), cams_i, [lsp[0][i:i + 1, :n1 + 1, :n2 + 1]], ass_i, pro_i, 0, f"plane-list pair {i} ({n1}x{n2})")


@pytest.mark.parametrize("u8", [False, True])
def test_resnet50_backbone_glue_on_host_matches_oracle(host_ops, u8):
    """Row f2: `ResNet50Backbone.forward` (stem normalisation + im2col, 53 convolutions as GEMMs on hi/lo planes with FrozenBN
    folded, max-pool, stride-2 paths, residual adds) on CPU tensors against the backbone oracle (= torchvision's resnet50,
    tests/test_oracle_backbone.py) at a reduced input size; parameter names / shapes are detectron2's."""
    from nopesac_b200 import backbone, config
    from oracle import backbone_restate as br
    cfg = config.inference_cfg(device="cpu")
    net = backbone.build_backbone(cfg)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == br.state_shapes()
    g = torch.Generator().manual_seed(8)
    sd = {}
    for k, shape in br.state_shapes().items():
        if k.endswith("running_var"):
            sd[k] = torch.rand(shape, generator=g) + 0.5
        elif k.endswith("norm.weight"):
            sd[k] = torch.rand(shape, generator=g) * 0.5 + 0.5
        elif k.endswith("norm.bias") or k.endswith("running_mean"):
            sd[k] = torch.randn(shape, generator=g) * 0.1
        else:
            fan_out = shape[0] * shape[2] * shape[3]
            sd[k] = torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5
    net.load_state_dict(sd)
    images = torch.rand(2, 3, 64, 96, generator=g) * 255
    if u8:      # the reference loader's format: the fast stem (one exact fp16 plane, normalisation in the weights, exact borders)
        images = images.to(torch.uint8)
    got = net(images)
    with torch.no_grad():
        want = br.resnet50(sd, br.normalize(images.float(), cfg.MODEL.PIXEL_MEAN, cfg.MODEL.PIXEL_STD))
    assert list(got) == ["res2", "res3", "res4", "res5"]
    for k in want:
        assert got[k].shape == want[k].shape, k
        rel = util.maxdiff(got[k], want[k]) / float(want[k].abs().max())
        assert rel <= 1e-4, (k, rel)
    assert {k: (v.channels, v.stride) for k, v in net.output_shape().items()} == {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)}


def test_backbone_entry_equals_python_stages_on_host(host_ops, monkeypatch):
    """nsac_backbone_forward (csrc/forward.cu compiled for the host, tensor engine = stand-ins) == the Python loop of
    backbone.py, bit for bit, on a small uint8 batch with odd sizes."""
    from nopesac_b200 import backbone, config, synthetic
    images = synthetic.make_images(3, 1, 37, 52)
    outs = {}
    for flag in (True, False):
        if flag:
            monkeypatch.delenv("NSAC_PY_STAGES", raising=False)
        else:
            monkeypatch.setenv("NSAC_PY_STAGES", "1")
        net = backbone.build_backbone(config.inference_cfg())
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        net.load_state_dict(synthetic.make_backbone_weights(shapes, seed=8))
        assert net.use_stage_entry == flag
        outs[flag] = net(images, planes=True)
    for k in ("res2", "res3", "res4", "res5"):
        (a, ha, wa), (b, hb, wb) = outs[True][k], outs[False][k]
        assert (ha, wa) == (hb, wb) and torch.equal(a.hi, b.hi) and torch.equal(a.lo, b.lo), k


def test_empty_batch_is_refused_loudly(host_ops):
    """B = 0 never reaches a kernel: the single-call entry refuses it with a message (the reference only ever runs one pair)."""
    head, match, _, _ = util.build_cuda_heads(32, "soft", 0.2, "cpu")
    z = lambda *s: torch.zeros(*s)
    with pytest.raises(RuntimeError, match="bad sizes B=0"):
        head(None, None, z(0, 8, 3), z(0, 8, 3), z(0, 8, 256), z(0, 8, 256), matching_net=match, initial_pose=(z(0, 3), z(0, 4)))
