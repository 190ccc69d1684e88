"""Pins the CPU oracle (oracle/ldconv_oracle.c) against fixtures minted from the reference's own LDConv
(/root/reference/ultralytics/nn/modules/conv.py:350-503) by oracle/gen_golden.py.  CPU only."""
import numpy as np
import pytest

from oracle import oracle
from tests import _golden

CASES = _golden.case_names()


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_fixture_set_complete():
    assert len(CASES) >= 14


@pytest.mark.parametrize("N", range(1, 13))
def test_p_n_table(N):
    """conv.py:413-432; SURVEY.md Appendix C.2 spells out N=2,3,5,7."""
    pn = oracle.p_n(N).reshape(-1)
    assert pn.dtype == np.int64 and pn.shape == (2 * N,)
    base = round(N ** 0.5)
    rows = [i // base for i in range(N)]
    cols = [i % base for i in range(N)]
    assert pn[:N].tolist() == rows and pn[N:].tolist() == cols
    known = {2: ([0, 1], [0, 0]), 3: ([0, 0, 1], [0, 1, 0]), 5: ([0, 0, 1, 1, 2], [0, 1, 0, 1, 0]),
             7: ([0, 0, 0, 1, 1, 1, 2], [0, 1, 2, 0, 1, 2, 0])}
    if N in known:
        assert (pn[:N].tolist(), pn[N:].tolist()) == known[N]


@pytest.mark.parametrize("name", CASES)
def test_p_n_matches_reference_buffer(name):
    z, prm, m = _golden.load(name)
    assert np.array_equal(oracle.p_n(m["N"]), z["param_p_n"])


@pytest.mark.parametrize("name", CASES)
def test_offset_conv(name):
    z, prm, m = _golden.load(name)
    off = oracle.offset_conv(z["x"], prm.p_conv_weight, prm.p_conv_bias, m["N"], m["s"])
    assert off.shape == z["offset"].shape
    np.testing.assert_allclose(off, z["offset"], rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(z["x"]).max())))


@pytest.mark.parametrize("name", CASES)
def test_indices_and_coords_bit_exact(name):
    """Given the reference's own offset tensor, corner indices and clamped coordinates are bit-exact."""
    z, prm, m = _golden.load(name)
    idx, coord, _ = oracle.grid(z["offset"], m["H"], m["W"], m["N"], m["s"])
    assert np.array_equal(idx, z["idx"])
    assert np.array_equal(coord.view(np.uint32), z["coord"].view(np.uint32))


@pytest.mark.parametrize("name", CASES)
def test_resampled_operand_bit_exact(name):
    """x_offset (what the reference feeds its (N,1) conv) is bit-exact given the reference's offsets."""
    z, prm, m = _golden.load(name)
    xo = oracle.sample(z["x"], z["offset"], m["N"], m["s"])
    assert np.array_equal(xo.view(np.uint32), z["x_offset"].view(np.uint32))


@pytest.mark.parametrize("name", CASES)
def test_forward_eval(name):
    z, prm, m = _golden.load(name)
    f = oracle.forward(z["x"], prm, training=False, offset=z["offset"])
    scale = max(1.0, float(np.abs(z["out_eval"]).max()))
    assert np.abs(f["out"] - z["out_eval"]).max() <= 1e-5 * scale
    # and end to end through the oracle's own offset conv
    f2 = oracle.forward(z["x"], prm, training=False)
    assert _rel(f2["out"], z["out_eval"]) <= 1e-4 or name.endswith("_far")


@pytest.mark.parametrize("name", CASES)
def test_forward_train_and_running_stats(name):
    z, prm, m = _golden.load(name)
    f = oracle.forward(z["x"], prm, training=True, offset=z["offset"])
    scale = max(1.0, float(np.abs(z["out_train"]).max()))
    assert np.abs(f["out"] - z["out_train"]).max() <= 2e-5 * scale
    np.testing.assert_allclose(prm.running_mean, z["train_running_mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(prm.running_var, z["train_running_var"], rtol=1e-5, atol=1e-6)
    assert int(z["train_num_batches_tracked"]) == 1


@pytest.mark.parametrize("name", CASES)
def test_backward_train(name):
    z, prm, m = _golden.load(name)
    f = oracle.forward(z["x"], prm, training=True, offset=z["offset"], update_running=False)
    g = oracle.backward(z["x"], prm, f, z["grad_out"], training=True)
    tol = 2e-4
    assert _rel(g["x"], z["train_grad_x"]) <= tol
    if name.endswith("_zero"):
        # one input channel and a BN in training mode: the output is invariant to the scale of conv.0.weight, so its
        # gradient (and everything downstream of it) is analytically zero and the reference's value is rounding noise
        return
    assert _rel(g["conv.0.weight"], z["train_grad_conv_0_weight"]) <= tol
    assert _rel(g["conv.1.weight"], z["train_grad_conv_1_weight"]) <= tol
    assert _rel(g["conv.1.bias"], z["train_grad_conv_1_bias"]) <= tol
    assert _rel(g["p_conv.weight"], z["train_grad_p_conv_weight"]) <= tol
    assert _rel(g["p_conv.bias"], z["train_grad_p_conv_bias"]) <= tol


@pytest.mark.parametrize("name", CASES)
def test_backward_eval(name):
    z, prm, m = _golden.load(name)
    f = oracle.forward(z["x"], prm, training=False, offset=z["offset"])
    g = oracle.backward(z["x"], prm, f, z["grad_out"], training=False)
    assert _rel(g["x"], z["eval_grad_x"]) <= 2e-4
    assert _rel(g["conv.0.weight"], z["eval_grad_conv0_weight"]) <= 2e-4
    assert _rel(g["p_conv.weight"], z["eval_grad_p_conv_weight"]) <= 2e-4


def test_quirk_last_row_doubled_at_zero_offset():
    """x = 1..16 on 4x4, N=1, s=1, zero offsets -> last row 26,28,30 and corner 64."""
    x = np.arange(1, 17, dtype=np.float32).reshape(1, 1, 4, 4)
    off = np.zeros((1, 2, 4, 4), dtype=np.float32)
    xo = oracle.sample(x, off, 1, 1)[0, 0]
    assert xo[3, :3].tolist() == [26.0, 28.0, 30.0] and xo[3, 3] == 64.0
    assert xo[:3, 3].tolist() == [8.0, 16.0, 24.0]
    assert np.array_equal(xo[:3, :3], x[0, 0, :3, :3])
    z, _, m = _golden.load("n1s1_zero")
    assert np.array_equal(oracle.sample(z["x"], z["offset"], 1, 1), z["x_offset"])


def test_quirk_stride2_edge_samples():
    """6x6 arange, N=3, s=2, zero offsets -> n=1 last column 12/36/60, n=2 last row 62/66/70."""
    x = np.arange(1, 37, dtype=np.float32).reshape(1, 1, 6, 6)
    off = np.zeros((1, 6, 3, 3), dtype=np.float32)
    xo = oracle.sample(x, off, 3, 2)[0, 0].reshape(3, 3, 3)      # (i, n, j)
    assert xo[:, 1, 2].tolist() == [12.0, 36.0, 60.0]
    assert xo[2, 2, :].tolist() == [62.0, 66.0, 70.0]
