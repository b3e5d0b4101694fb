"""CPU: the backbone oracle (oracle/backbone_restate.py, row f2) pinned against torchvision's resnet50 — the network whose
weights the reference loads (`detectron2://ImageNetPretrained/torchvision/R-50.pkl`, Base.yaml:5) — stage by stage, bit for bit."""
import torch

from oracle import backbone_restate as br


def _seeded_torchvision():
    import torchvision
    torch.manual_seed(0)
    m = torchvision.models.resnet50(weights=None).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, b in m.named_buffers():            # non-trivial running statistics
            if name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif name.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
        for name, p in m.named_parameters():
            if "bn" in name or "downsample.1" in name:
                p.copy_(torch.rand(p.shape, generator=g) + 0.5 if name.endswith("weight") else torch.randn(p.shape, generator=g) * 0.1)
    return m


def test_backbone_oracle_equals_torchvision_resnet50():
    m = _seeded_torchvision()
    sd = br.from_torchvision(m.state_dict())
    assert {k: tuple(v.shape) for k, v in sd.items()} == br.state_shapes()
    x = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = {}
        y = m.maxpool(m.relu(m.bn1(m.conv1(x))))
        for name, layer in (("res2", m.layer1), ("res3", m.layer2), ("res4", m.layer3), ("res5", m.layer4)):
            y = layer(y)
            want[name] = y
        got = br.resnet50(sd, x)
    for k in want:
        assert got[k].shape == want[k].shape
        # FrozenBN applies scale / bias in a different association than BatchNorm's fused kernel: equal to fp32 rounding
        assert float((got[k] - want[k]).abs().max()) <= 2e-5 * float(want[k].abs().max()), k
    assert got["res5"].shape == (2, 2048, 2, 3)


def test_normalize_matches_reference_constants():
    img = torch.rand(1, 3, 4, 4) * 255
    out = br.normalize(img, [123.675, 116.280, 103.530], [58.395, 57.120, 57.375])
    assert abs(float(out[0, 1, 0, 0]) - (float(img[0, 1, 0, 0]) - 116.28) / 57.12) <= 1e-6
