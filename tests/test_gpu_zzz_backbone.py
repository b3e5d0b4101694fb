"""GPU: the ResNet-50 backbone (row f2) on the tensor-core engine against the backbone oracle (= torchvision's resnet50,
tests/test_oracle_backbone.py), and RGB -> backbone -> camera head.  Named to run last: the backbone's glue kernels and its
whole Python composition were executed on the host against the same oracle (tests/test_simt_host_kernels.py,
tests/test_host_head_glue.py); these tests put the real tcgen05 GEMM / implicit-GEMM convolution underneath.
Bar: 1e-4 of the stage's largest activation (3-pass fp16 hi/lo planes ~ fp32 through 53 convolutions)."""
import pytest
import torch

from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]   # code that has not met a GPU yet: never hang the run


def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _seeded_state(seed=8):
    from oracle import backbone_restate as br
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in br.state_shapes().items():
        if k.endswith("running_var"):
            sd[k] = torch.rand(shape, generator=g) + 0.5
        elif k.endswith("norm.weight"):
            sd[k] = torch.rand(shape, generator=g) * 0.5 + 0.5
        elif k.endswith("norm.bias") or k.endswith("running_mean"):
            sd[k] = torch.randn(shape, generator=g) * 0.1
        else:
            sd[k] = torch.randn(shape, generator=g) * (2.0 / (shape[0] * shape[2] * shape[3])) ** 0.5
    return sd


@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("N,H,W", [(2, 64, 96), (1, 480, 640), (3, 120, 200), (2, 97, 131)])
def test_resnet50_backbone_matches_oracle(N, H, W, u8):
    dev = _gpu()
    from nopesac_b200 import backbone, config
    from oracle import backbone_restate as br
    cfg = config.inference_cfg()
    net = backbone.build_backbone(cfg)
    sd = _seeded_state()
    net.load_state_dict(sd)
    net = net.to(dev)
    images = torch.rand(N, 3, H, W, generator=torch.Generator().manual_seed(N + H)) * 255
    if u8:          # the reference loader's format -> the fast stem (one exact fp16 plane + exact border recomputation)
        images = images.to(torch.uint8)
    got = net(images.to(dev))
    torch.cuda.synchronize()
    with torch.no_grad():
        want = br.resnet50(sd, br.normalize(images.float(), cfg.MODEL.PIXEL_MEAN, cfg.MODEL.PIXEL_STD))
    for k in want:
        assert got[k].shape == want[k].shape, k
        rel = util.maxdiff(got[k], want[k]) / float(want[k].abs().max())
        assert rel <= 1e-4, (k, rel)


def test_rgb_to_camera_head():
    """`PlaneTR_NopeSAC(cfg, with_backbone=True).inference_from_images` from uint8 RGB, at the north-star bars:
    (1) the planes hand-off (backbone -> PlaneFeatures -> pixel network, no NCHW round trip) == the same head called the
        reference way with this backbone's feature maps as fp32 NCHW dicts: assignments EXACT, every pose 1e-4;
    (2) those fp32 feature maps through the CPU oracle head == the CUDA head: assignments / matched_num exact, poses 1e-4 —
        K1 + matcher + refinement parity on backbone-produced (not synthetic) features;
    the backbone itself is held to the oracle in test_resnet50_backbone_matches_oracle."""
    dev = _gpu()
    from nopesac_b200 import config, meta_arch, synthetic
    from oracle import restate
    NQ = 50
    cfg = config.inference_cfg(NQ)
    model = meta_arch.PlaneTR_NopeSAC(cfg, with_backbone=True)
    sd, msd = util.make_weights(NQ)
    model.camera_head_list[0].load_state_dict(sd)
    model.matching_head.load_state_dict(msd)
    shapes = {k: tuple(v.shape) for k, v in model.backbone.state_dict().items()}
    model.backbone.load_state_dict(synthetic.make_backbone_weights(shapes, seed=9))
    model = model.to(dev)
    head = model.camera_head_list[0]
    B = 2
    b = synthetic.make_batch(7, B, 16)
    images = synthetic.make_images(5, 2 * B)                    # uint8 [2B,3,480,640]: first views, then second views
    bd = b.to(dev)
    got = model.inference_from_images(images.to(dev), None, bd.planes1, bd.planes2, bd.app1, bd.app2)
    feats = model.backbone(images.to(dev))                      # fp32 NCHW, the reference's backbone interface
    f1 = {k: v[:B].contiguous() for k, v in feats.items()}
    f2 = {k: v[B:].contiguous() for k, v in feats.items()}
    want = head(f1, f2, bd.planes1, bd.planes2, planeApp1=bd.app1, planeApp2=bd.app2, matching_net=model.matching_head)
    torch.cuda.synchronize()
    for key in ("pred_assignment_beforeRef0", "pred_assignment"):
        assert torch.equal(got[4][key], want[4][key]), key
    assert torch.equal(got[5]["matched_num"], want[5]["matched_num"])
    for key in ("camera_init", "camera_initRec", "camera_avgRef0", "camera"):
        assert util.maxdiff(got[0][key]["tran"], want[0][key]["tran"]) <= util.ABS_TOL, key
        assert util.maxdiff(got[0][key]["rot"], want[0][key]["rot"]) <= util.ABS_TOL, key
    # (2) the oracle head on the same feature maps
    for i in range(B):
        with torch.no_grad():
            o = restate.inference_joint(sd, msd, {k: v[i:i + 1].cpu() for k, v in f1.items()}, {k: v[i:i + 1].cpu() for k, v in f2.items()},
                                        b.planes1[i:i + 1], b.planes2[i:i + 1], b.app1[i:i + 1], b.app2[i:i + 1], num_queries=NQ)
        assert int(want[5]["matched_num"][i]) == o["matched_num"], i
        assert torch.equal(want[4]["pred_assignment_beforeRef0"][i].cpu(), o["assignment_before"][0]), i
        assert torch.equal(want[4]["pred_assignment"][i].cpu(), o["assignment_after"][0]), i
        for key, ok in (("camera_init", "camera_init"), ("camera_initRec", "camera_initRec"), ("camera", "camera")):
            assert util.maxdiff(want[0][key]["tran"][i], o[ok][0][0]) <= util.ABS_TOL, (key, i, util.maxdiff(want[0][key]["tran"][i], o[ok][0][0]))
            assert util.maxdiff(want[0][key]["rot"][i], o[ok][1][0]) <= util.ABS_TOL, (key, i, util.maxdiff(want[0][key]["rot"][i], o[ok][1][0]))


def test_s5_full_size_properties():
    """Size-independent properties of the headline configuration (BASELINE configs[1]: 16 planes x 256 hypotheses, NQ = 256, from
    uint8 RGB) that need no oracle at full size:
      * determinism — two runs of the same batch give identical bytes (atomics-free reductions, fixed summation orders);
      * batch invariance — pairs are independent: a pair computed alone, in a batch of 6, or at another position of a permuted
        batch has the same result rows (1e-6: tiles are per image / per pair, so the arithmetic of a row does not depend on its
        neighbours), with identical assignment matrices and matched counts;
      * the result rows are well formed: unit quaternions, matched_num = 256, finite."""
    dev = _gpu()
    from nopesac_b200 import config, meta_arch, synthetic
    NQ, P, B = 256, 16, 6
    model = meta_arch.PlaneTR_NopeSAC(config.inference_cfg(NQ), with_backbone=True)
    sd, msd = util.make_weights(NQ)
    model.camera_head_list[0].load_state_dict(sd)
    model.matching_head.load_state_dict(msd)
    shapes = {k: tuple(v.shape) for k, v in model.backbone.state_dict().items()}
    model.backbone.load_state_dict(synthetic.make_backbone_weights(shapes, seed=8))
    model = model.to(dev)
    hp = synthetic.all_pairs_hypotheses(P, NQ).to(dev, torch.int32)
    b = synthetic.make_batch(9000, B, P).to(dev)
    images = synthetic.make_images(9000, 2 * B).to(dev)

    def run(idx):
        idx = torch.as_tensor(idx, device=dev)
        img = torch.cat([images[idx], images[B + idx]])
        out = model.inference_from_images(img, None, b.planes1[idx], b.planes2[idx], b.app1[idx], b.app2[idx], hyp_pairs=hp)
        torch.cuda.synchronize()
        return out[5]["pose"].clone(), out[4]["pred_assignment"].clone(), out[5]["matched_num"].clone(), out[0]["camera_init"]["rot"].clone()

    pose, ass, m, q_init = run(list(range(B)))
    pose2, ass2, m2, _ = run(list(range(B)))
    assert torch.equal(pose, pose2) and torch.equal(ass, ass2) and torch.equal(m, m2), "two runs of the same batch differ"
    assert torch.isfinite(pose).all() and m.cpu().tolist() == [NQ] * B
    for cols in (slice(3, 7), slice(10, 14)):
        assert util.maxdiff(pose[:, cols].norm(dim=-1), torch.ones(B)) <= 1e-5
    assert bool((q_init[:, 0] >= 0).all())
    perm = [4, 0, 5, 2, 1, 3]
    pose_p, ass_p, m_p, _ = run(perm)
    assert util.maxdiff(pose_p, pose[perm]) <= 1e-6, util.maxdiff(pose_p, pose[perm])
    assert torch.equal(ass_p, ass[perm]) and torch.equal(m_p, m[perm])
    for i in (0, 3):
        pose_1, ass_1, m_1, _ = run([i])
        assert util.maxdiff(pose_1, pose[i:i + 1]) <= 1e-6, (i, util.maxdiff(pose_1, pose[i:i + 1]))
        assert torch.equal(ass_1, ass[i:i + 1]) and torch.equal(m_1, m[i:i + 1])


@pytest.mark.parametrize("H,W", [(70, 90), (33, 47), (480, 640)])
def test_stem_uint8_paths_match_reference_convolution(H, W):
    """The two uint8 stem routes against relu(conv2d(normalise(img), w, b, stride 2, padding 3)) in fp64, border pixels included:
    (A) border-class indicator columns in the im2col matrix + correction rows in the weight matrix (the path the backbone takes),
    (B) plain im2col + exact fp32 recomputation of the border pixels (nsac_stem_border_fix, kept for tiny images)."""
    dev = _gpu()
    import torch.nn.functional as F
    from nopesac_b200 import backbone, config, ops
    cfg = config.inference_cfg()
    net = backbone.build_backbone(cfg)
    net.load_state_dict(_seeded_state())
    net = net.to(dev)
    pk = net.prepare()
    N = 2
    images = torch.randint(0, 256, (N, 3, H, W), generator=torch.Generator().manual_seed(H + W), dtype=torch.uint8)
    wf, bf = net.stem.conv1.folded()                                  # [64, 147] in (ky, kx, c) order, FrozenBN folded
    w4 = wf.view(64, 7, 7, 3).permute(0, 3, 1, 2).double().cpu()
    mean = torch.tensor(cfg.MODEL.PIXEL_MEAN, dtype=torch.float64).view(1, 3, 1, 1)
    std = torch.tensor(cfg.MODEL.PIXEL_STD, dtype=torch.float64).view(1, 3, 1, 1)
    ref = F.relu(F.conv2d((images.double() - mean) / std, w4, bf.double().cpu(), stride=2, padding=3))
    ref = ref.permute(0, 2, 3, 1).reshape(-1, 64)
    scale = float(ref.abs().max())
    img = images.to(dev)
    cols, Ho, Wo = ops.stem_im2col_u8(img, border_classes=True)
    ws, bs = net._stem_u8_weights(pk, H, W)
    xa, _ = ops.gemm_tc(cols, ws, bs, ops.ACT_RELU)
    cols_b, _, _ = ops.stem_im2col_u8(img)
    wsb, bsb = pk["stem.u8"]
    xb, _ = ops.gemm_tc(cols_b, wsb, bsb, ops.ACT_RELU)
    wf32, bf32 = pk["stem.f32"]
    ops.stem_border_fix(img, wf32, bf32, net.pixel_mean, net.pixel_std, xb)
    torch.cuda.synchronize()
    assert (Ho, Wo) == ((H - 1) // 2 + 1, (W - 1) // 2 + 1) and xa.shape == ref.shape
    for name, x in (("border classes", xa), ("border recomputation", xb)):
        d = (x.double().cpu() - ref).abs().view(N, Ho, Wo, 64)
        assert float(d.max()) <= 3e-6 * scale, (name, float(d.max()) / scale)
        ring = torch.cat([d[:, :2].reshape(-1), d[:, -2:].reshape(-1), d[:, :, :2].reshape(-1), d[:, :, -2:].reshape(-1)])
        assert float(ring.max()) <= 3e-6 * scale, (name, "border ring", float(ring.max()) / scale)
