"""The numpy restatement of the reference's non_max_suppression / soft_nms (oracle/nms_oracle.py) against the golden vectors
minted from the reference function itself (oracle/gen_nms_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import nms_oracle
from tests import _golden

CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(_golden.GOLDEN_DIR, "nms_*.npz")))


def load(name):
    z = np.load(os.path.join(_golden.GOLDEN_DIR, f"nms_{name}.npz"))
    kw = dict(conf_thres=float(z["conf_thres"]), iou_thres=float(z["iou_thres"]), agnostic=bool(z["agnostic"]), max_det=int(z["max_det"]))
    want = [z[f"out{i}"] for i in range(len(z["counts"]))]
    return z["pred"], kw, want


@pytest.mark.parametrize("name", CASES)
def test_nms_oracle_matches_reference_fixture(name):
    pred, kw, want = load(name)
    got = nms_oracle.non_max_suppression(pred, **kw)
    assert [g.shape[0] for g in got] == [w.shape[0] for w in want]
    for g, w in zip(got, want):
        assert np.array_equal(g[:, 5], w[:, 5])                      # classes
        assert np.array_equal(g[:, :4], w[:, :4])                    # boxes: the same fp32 arithmetic
        assert np.abs(g[:, 4] - w[:, 4]).max(initial=0.0) <= 2e-7    # decayed confidences (exp implementations differ by an ulp)


def test_fixtures_cover_the_quirks():
    assert {"default", "bf16_agnostic", "many_small_maxdet", "lowconf", "edge_counts"} <= set(CASES)
    _, _, want = load("edge_counts")
    assert [w.shape[0] for w in want] == [0, 0, 1]      # no candidate / one candidate is dropped / two candidates keep one
