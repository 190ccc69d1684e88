"""The eager CPU port that bench.py times as `cpu_baseline` (oracle/ldconv_torch_port.py) against the golden vectors
minted from the reference LDConv, and the benchmark graph (experiment_yolo_b200/dealyolo.py) against the golden output of
the reference DetectionModel.  CPU only."""
import os

import numpy as np
import pytest
import torch

from experiment_yolo_b200 import dealyolo
from oracle.ldconv_torch_port import LDConvTorchPort
from tests import _golden

CASES = _golden.case_names()


def _port_from_golden(z, prm, m):
    mod = LDConvTorchPort(m["inc"], m["outc"], m["N"], m["s"])
    sd = {k: torch.from_numpy(np.array(z["param_" + k.replace(".", "_")])) for k in mod.state_dict()}
    mod.load_state_dict(sd, strict=True)
    mod.conv[1].eps, mod.conv[1].momentum = prm.eps, prm.momentum
    return mod


@pytest.mark.parametrize("name", CASES)
def test_port_forward_bit_exact(name):
    torch.set_num_threads(1)
    z, prm, m = _golden.load(name)
    mod = _port_from_golden(z, prm, m).eval()
    with torch.no_grad():
        y = mod(torch.from_numpy(z["x"]))
    assert np.array_equal(y.numpy(), z["out_eval"])


@pytest.mark.parametrize("name", CASES)
def test_port_backward(name):
    torch.set_num_threads(1)
    z, prm, m = _golden.load(name)
    mod = _port_from_golden(z, prm, m).train()
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    y = mod(x)
    y.backward(torch.from_numpy(z["grad_out"]))
    np.testing.assert_allclose(y.detach().numpy(), z["out_train"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), z["train_grad_x"], rtol=1e-4, atol=1e-5)


def test_model_graph_matches_reference_detection_model():
    """DealYolo + the port reproduce the reference DetectionModel('yolov8-LD-P2.yaml') output stored by
    oracle/gen_model_golden.py (which asserted bit-equality against the live reference when it ran)."""
    z = np.load(os.path.join(_golden.GOLDEN_DIR, "model_deal_yolo_ld.npz"))
    model = dealyolo.DealYolo(nc=6, ldconv_cls=LDConvTorchPort)
    assert sum(p.numel() for p in model.parameters()) == int(z["n_params"]) == 918304      # README.md:61 "0.914 M"
    assert [float(s) for s in model.stride] == z["strides"].tolist() == [4.0, 8.0, 16.0]
    model.load_state_dict(dealyolo.seeded_state(model, seed=0), strict=True)
    model.eval()
    with torch.no_grad():
        y, feats = model(torch.from_numpy(z["x"]))
    assert tuple(y.shape) == (1, 10, 1008)
    np.testing.assert_allclose(y.numpy(), z["y"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(feats[0].numpy(), z["feat0"], rtol=1e-5, atol=1e-5)


def test_model_yaml_hook_resolves_ldconv_rows_by_name():
    model = dealyolo.DealYolo(nc=6, ldconv_cls=LDConvTorchPort)
    ld = model.ldconv_layers()
    assert [m.i for m in ld] == [0, 1, 3, 5, 8, 10, 13, 15, 18, 21]
    assert [(m.p_conv.in_channels, m.conv[0].out_channels, m.num_param, m.stride) for m in ld] == [
        (3, 16, 3, 2), (16, 32, 3, 2), (32, 64, 3, 2), (64, 128, 3, 2), (128, 64, 1, 1), (64, 64, 1, 1), (64, 32, 1, 1),
        (32, 32, 1, 1), (32, 32, 3, 2), (64, 64, 3, 2)]
    assert all(m.conv[1].eps == 1e-3 and m.conv[1].momentum == 0.03 for m in ld)           # torch_utils.py:342-352


def test_strides_agree_with_a_probe_forward():
    model = dealyolo.DealYolo(nc=6, ldconv_cls=LDConvTorchPort).train()
    with torch.no_grad():
        feats = model(torch.zeros(2, 3, 64, 64))
    assert [64 / f.shape[-2] for f in feats] == [4.0, 8.0, 16.0]


def test_port_graph_reproduces_the_640_fixture_of_the_reference_model():
    """BASELINE config 1 at its real size: the benchmark graph + eager port on the CPU equals the fixture minted from the
    reference DetectionModel (oracle/gen_model_golden.py; bit for bit there, with the generator's thread count -- the CPU
    conv kernels partition by thread count, so here: boxes to 1e-3 of 640 pixels, scores to 1e-6) -- the fixture the GPU tests of
    tests/test_gpu_fullsize.py compare with, and the model `--impl reference` / `cpu_baseline` time."""
    import os
    z = np.load(os.path.join(_golden.GOLDEN_DIR, "model_deal_yolo_ld_640.npz"))
    x = torch.rand(1, 3, 640, 640, generator=torch.Generator().manual_seed(0))
    assert np.array_equal(x.reshape(-1)[::4099].numpy(), z["x_probe"])
    model = dealyolo.DealYolo(nc=6, ldconv_cls=LDConvTorchPort)
    model.load_state_dict(dealyolo.seeded_state(model, seed=0), strict=True)
    model.eval()
    with torch.no_grad():
        y, _ = model(x)
    d = np.abs(y.numpy() - z["y"])
    assert d[:, :4].max() <= 1e-3 and d[:, 4:].max() <= 1e-6
