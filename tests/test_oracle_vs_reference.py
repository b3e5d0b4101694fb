"""CPU, build container only: the oracle restatement against the LIVE reference code imported from
/root/reference (skipped where the reference tree is absent, e.g. on the GPU box)."""
import pytest
import torch

from oracle import ref_loader, restate
from nopesac_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.mark.parametrize("nq,planes", [(32, 8), (50, 16)])
def test_inference_joint_bit_exact_vs_reference(nq, planes):
    head, match, _ = ref_loader.build_heads(num_queries=nq)       # reference initialisers, seed 40
    sd = {k: v.detach() for k, v in head.state_dict().items()}
    msd = {k: v.detach() for k, v in match.state_dict().items()}
    for idx in range(2):
        b = synthetic.make_batch(idx, 1, planes, with_features=True)
        with ref_loader.cpu_patch(), torch.no_grad():
            cams, _, _, lsp, ass, pro = head(b.feats1, b.feats2, b.planes1, b.planes2, b.app1, b.app2, matching_net=match)
        with torch.no_grad():
            o = restate.inference_joint(sd, msd, b.feats1, b.feats2, b.planes1, b.planes2, b.app1, b.app2, num_queries=nq)
        assert torch.equal(lsp[0], o["log_scores_padded"])
        assert torch.equal(ass["pred_assignment_beforeRef0"], o["assignment_before"])
        assert torch.equal(ass["pred_assignment"], o["assignment_after"])
        for key, mine in (("camera_init", o["camera_init"]), ("camera_initRec", o["camera_initRec"]),
                          ("camera_avgRef0", o["camera_avgRef0"]), ("camera", o["camera"])):
            assert torch.equal(cams[key]["tran"], mine[0]) and torch.equal(cams[key]["rot"], mine[1]), key
        assert torch.equal(pro["score_soft_rot"], o["ref"]["score_soft_rot"])
        assert torch.equal(pro["sig_seq"], o["sig_seq"][:o["matched_num"], 0])


def test_cpu_path_of_reference_needs_the_cuda_patch():
    """MODEL.DEVICE=cpu does not run as shipped (matching_head.py:274-301 hard-codes .cuda())."""
    head, match, _ = ref_loader.build_heads(num_queries=32)
    b = synthetic.make_batch(0, 1, 8)
    cam = torch.tensor([[0., 0., 0., 1., 0., 0., 0.]])
    with pytest.raises((RuntimeError, AssertionError)):
        match(b.app1, b.app2, cam, b.planes1, b.planes2)
