"""CPU: the oracle (oracle/restate.py) reproduces every golden fixture that the LIVE reference produced
(tests/golden/make_golden.py).  Same torch ops in the same order on the same machine class -> the bar is
1e-6 abs (bit-exact where BLAS kernels agree), with index/selection outputs exactly equal."""
import pytest
import torch

from tests import util

EXACT = ("assignment_before", "assignment_after", "sig_seq", "matched_num")


@pytest.mark.parametrize("name", util.golden_names())
def test_oracle_matches_golden(name):
    g = util.load_golden(name)
    case = g["case"]
    sd, msd = util.make_weights(case["NQ"], g["head_seed"], g["match_seed"])
    for pair_idx, want in zip(case["pairs"], g["outputs"]):
        got = util.oracle_to_flat(util.run_oracle_case(case, pair_idx, sd, msd))
        for k, w in want.items():
            assert k in got, f"{name}: oracle lacks {k}"
            gk = got[k].reshape(w.shape)
            if k in EXACT:
                assert torch.equal(gk.float(), w.float()), f"{name} pair {pair_idx}: {k} differs"
            elif k == "log_scores_padded":
                assert util.maxdiff(gk.exp(), w.exp()) <= 1e-6
                assert float(((gk - w).abs() / w.abs().clamp_min(1.0)).max()) <= 1e-5
            else:
                assert util.maxdiff(gk, w) <= 2e-6, f"{name} pair {pair_idx}: {k} off by {util.maxdiff(gk, w)}"
