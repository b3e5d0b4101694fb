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
