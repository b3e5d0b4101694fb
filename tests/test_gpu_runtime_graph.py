"""GPU: the CUDA-graph runner replays exactly the eager forward (same kernels, same inputs -> identical result rows),
and follows new inputs loaded into its static buffers."""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def test_graphed_head_equals_eager():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    dev = torch.device("cuda:0")
    from nopesac_b200 import synthetic
    from nopesac_b200.runtime import GraphedCameraHead
    NQ, P, B = 32, 8, 3
    head, match, _, _ = util.build_cuda_heads(NQ, "soft", 0.2, dev)

    def device_batch(first):
        b = synthetic.make_batch(first, B, P, with_features=True).to(dev)
        return {"planes1": b.planes1, "planes2": b.planes2, "app1": b.app1, "app2": b.app2, "feats1": b.feats1, "feats2": b.feats2}

    def eager(d):
        return head(d["feats1"], d["feats2"], d["planes1"], d["planes2"], d["app1"], d["app2"], matching_net=match)[5]["pose"].clone()

    b0, b1 = device_batch(0), device_batch(100)
    want0, want1 = eager(b0), eager(b1)
    runner = GraphedCameraHead(head, match, {k: (dict(v) if isinstance(v, dict) else v) for k, v in b0.items()})
    got0 = runner().clone()
    runner.load(b1)
    got1 = runner().clone()
    torch.cuda.synchronize()
    assert torch.equal(got0, want0) and torch.equal(got1, want1)
    assert not torch.equal(want0, want1)
