"""GPU parity at BASELINE.json's REAL sizes: every LDConv row of yolov8-LD-P2 at its 640x640 map
size (and layers 0 / 1 at 1280x1280, config 5) against the CPU oracle, batch invariance of the persistent kernels at
batch 64, the full model at 1x3x640x640 (config 1) against the fixture minted from the reference DetectionModel, and the
measured index-flip rate of the one fixture whose end-to-end comparison is otherwise skipped.

Reference path: /root/reference/ultralytics/nn/modules/conv.py:366-410 (LDConv.forward); bars as in test_gpu_parity.py:
indices bit-exact given the offsets, bf16 outputs rel-L2 <= 1e-2 against the fp32 oracle on bf16-rounded tensors.
"""
import os

import numpy as np
import pytest
import torch

import experiment_yolo_b200 as E
from experiment_yolo_b200 import _lib, dealyolo
from oracle import oracle
from tests import _golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (yaml row, C, O, num_param, stride, input map size at 640x640): SURVEY.md Appendix B
YAML_LAYERS = [(0, 3, 16, 3, 2, 640), (1, 16, 32, 3, 2, 320), (3, 32, 64, 3, 2, 160), (5, 64, 128, 3, 2, 80),
               (8, 128, 64, 1, 1, 40), (10, 64, 64, 1, 1, 80), (13, 64, 32, 1, 1, 80), (15, 32, 32, 1, 1, 160),
               (18, 32, 32, 3, 2, 160), (21, 64, 64, 3, 2, 80)]


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _rnd(t):
    return t.detach().bfloat16().float().numpy()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _make_module(C, O, N, s, sigma, bias_sigma=None, seed=0):
    torch.manual_seed(seed * 1000 + C * 7 + O + N)
    mod = E.LDConv(C, O, N, s)
    with torch.no_grad():
        mod.p_conv.weight.normal_(0, sigma)
        if bias_sigma is not None:
            mod.p_conv.bias.normal_(0, bias_sigma)
        mod.conv[1].running_mean.normal_(0, 0.3)
        mod.conv[1].running_var.uniform_(0.5, 1.5)
        mod.conv[1].weight.uniform_(0.5, 1.5)
        mod.conv[1].bias.normal_(0, 0.2)
    mod.conv[1].eps, mod.conv[1].momentum = 1e-3, 0.03
    prm = oracle.LDConvParams(_rnd(mod.p_conv.weight), _rnd(mod.p_conv.bias), _rnd(mod.conv[0].weight), _rnd(mod.conv[1].weight),
                              _rnd(mod.conv[1].bias), _rnd(mod.conv[1].running_mean), _rnd(mod.conv[1].running_var), N, s,
                              1e-3, 0.03)
    return mod, prm


def _device_indices(x_np, off_np, N, s):
    """ldconv_gather_fwd's debug outputs (corner indices, clamped coordinates) for bf16 x and the given fp32 offsets"""
    L = _lib.load()
    B, C, H, W = x_np.shape
    h, w = (H - 1) // s + 1, (W - 1) // s + 1
    x = torch.from_numpy(np.ascontiguousarray(x_np.transpose(0, 2, 3, 1))).to(DEV).bfloat16()
    off = torch.from_numpy(np.ascontiguousarray(off_np.transpose(0, 2, 3, 1))).to(DEV)
    pn = torch.tensor(_lib.p_n_table(N), dtype=torch.int32, device=DEV)
    M = B * h * w
    operand = torch.empty((M, N * C), device=DEV, dtype=torch.bfloat16)
    idx = torch.empty((M, N, 4), device=DEV, dtype=torch.int32)
    coord = torch.empty((M, N, 2), device=DEV, dtype=torch.float32)
    _lib.check(L.ldconv_gather_fwd(_ptr(x), _ptr(off), _ptr(pn), _ptr(operand), _ptr(idx), _ptr(coord), B, C, H, W, N, s,
                                   _lib.BF16, _stream()), "ldconv_gather_fwd")
    torch.cuda.synchronize()
    return idx.cpu().numpy(), coord.cpu().numpy(), operand


def _check_layer(C, O, N, s, H, W, B, sigma, bias_sigma=None):
    mod, prm = _make_module(C, O, N, s, sigma, bias_sigma)
    x = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(C + H))
    x_np = _rnd(x)
    f = oracle.forward(x_np, prm, training=False)
    # (1) sampling indices / clamped coordinates bit-exact given the oracle's offsets, at the real map size
    idx, coord, _ = _device_indices(x_np, f["offset"], N, s)
    oi, oc, _ = oracle.grid(f["offset"], H, W, N, s)
    assert np.array_equal(idx.reshape(-1), oi.reshape(-1))
    assert np.array_equal(coord.view(np.uint32).reshape(-1), oc.view(np.uint32).reshape(-1))
    # (2) the inference path the benchmark step runs for this shape (one-kernel layer 0 / offset conv + gather+GEMM kernel)
    dmod = mod.to(DEV).bfloat16().eval()
    xd = torch.from_numpy(x_np).to(DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    _lib.call_counts.clear()
    with torch.no_grad():
        y = dmod(xd)
    torch.cuda.synchronize()
    used = set(_lib.call_counts)
    assert used & {"ldconv_fused_fwd", "ldconv_gather_gemm_fwd", "ldconv_onepass_fwd"}, used      # not the training fallback
    y = y.float().cpu().numpy()
    assert y.shape == f["out"].shape
    rel = _rel(y, f["out"])
    assert rel <= 1e-2, rel
    return rel


@pytest.mark.parametrize("sigma", [0.05, 0.3])
@pytest.mark.parametrize("layer,C,O,N,s,HW", YAML_LAYERS)
def test_yaml_layer_at_real_map_size_vs_oracle(layer, C, O, N, s, HW, sigma):
    """BASELINE config 3's LDConv shapes at their true map sizes (640x640 image), B = 2, two offset scales."""
    _check_layer(C, O, N, s, HW, HW, 2, sigma)


@pytest.mark.parametrize("layer,C,O,N,s,HW", [(0, 3, 16, 3, 2, 1280), (1, 16, 32, 3, 2, 640)])
@pytest.mark.parametrize("sigma,bias_sigma", [(0.05, 0.0), (0.3, 2.0), (0.05, 8.0)])
def test_config5_1280_layers_vs_oracle(layer, C, O, N, s, HW, sigma, bias_sigma):
    """BASELINE config 5 (1280x1280 tiles): the two largest maps of the model, with the offset statistics SURVEY.md 8d names
    (p_conv.weight sigma, bias sigma in pixels) to stress the staged halo and the clamp quirk."""
    _check_layer(C, O, N, s, HW, HW, 1, sigma, bias_sigma)


@pytest.mark.parametrize("layer,C,O,N,s,HW", [YAML_LAYERS[0], YAML_LAYERS[1], YAML_LAYERS[7], YAML_LAYERS[4], YAML_LAYERS[9]])
def test_batch_invariance_batch64_equals_32_runs_of_2(layer, C, O, N, s, HW):
    """Batch 64 (the benchmark's launch: persistent multi-tile loops, tiles >> CTAs, 32-bit byte offsets of the largest
    images) gives bit for bit what the same images give two at a time."""
    mod, _ = _make_module(C, O, N, s, 0.1, 1.0)
    dmod = mod.to(DEV).bfloat16().eval()
    g = torch.Generator(device=DEV).manual_seed(layer)
    x = torch.randn((64, C, HW, HW), device=DEV, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y64 = dmod(x)
        for b0 in range(0, 64, 2):
            y2 = dmod(x[b0:b0 + 2])
            assert torch.equal(y2, y64[b0:b0 + 2]), (layer, b0)


# ---------------------------------------------------------------------------------------------- config 1: 1x3x640x640 ----
def _load640():
    z = np.load(os.path.join(_golden.GOLDEN_DIR, "model_deal_yolo_ld_640.npz"))
    x = torch.rand(1, 3, 640, 640, generator=torch.Generator().manual_seed(0))
    assert abs(float(x.double().sum()) - float(z["x_sum"])) <= 1e-6 and np.array_equal(x.reshape(-1)[::4099].numpy(), z["x_probe"]), \
        "torch.rand(seed 0) no longer reproduces the fixture's input"
    model = dealyolo.DealYolo(nc=6)
    model.load_state_dict(dealyolo.seeded_state(model, seed=0), strict=True)
    return z, x, model


def test_config1_full_model_640_fp32_vs_reference_fixture():
    """BASELINE config 1: DEAL-YOLO-LD forward, batch 1, 640x640, fp32: the CUDA LDConv inside the graph on the GPU vs the
    reference DetectionModel on the CPU (fixture: oracle/gen_model_golden.py).  Boxes are in pixels (up to 640)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    z, x, model = _load640()
    model = model.to(DEV).eval()
    with torch.no_grad():
        y, feats = model(x.to(DEV))
    y, ref = y.cpu().numpy(), z["y"]
    assert y.shape == ref.shape == (1, 10, 33600)
    assert np.abs(y[:, 4:] - ref[:, 4:]).max() <= 1e-4            # class scores
    assert np.abs(y[:, :4] - ref[:, :4]).max() <= 1e-2            # boxes: 640-pixel scale (1.6e-5 relative)


def test_config1_full_model_640_bf16_fused_engine_vs_reference_fixture():
    """The same fixture through the bf16 executor the benchmark times (engine.FusedDealYolo)."""
    from experiment_yolo_b200 import engine
    z, x, model = _load640()
    model = dealyolo.channels_last_(model.to(DEV).bfloat16().eval())
    eng = engine.FusedDealYolo(model)
    y, _ = eng(x.to(DEV).bfloat16().contiguous(memory_format=torch.channels_last))
    torch.cuda.synchronize()
    y, ref = y.float().cpu().numpy(), z["y"]
    assert np.isfinite(y).all()
    rel = _rel(y, ref)
    assert rel <= 2e-2, rel
    assert np.abs(y[:, 4:] - ref[:, 4:]).max() <= 0.03


# ------------------------------------------------------------------------------- the skipped fixture, measured instead ----
def test_far_offset_fixture_index_flip_rate_is_small_and_rest_matches():
    """`n3s2_far` (offsets of ~8 px on a 10x14 image) is excluded from the end-to-end comparisons because the output is
    discontinuous in the offsets at p = H-1.  Here the exclusion is MEASURED: the bf16 tensor-core offset
    conv vs the oracle's fp32 offsets on the same bf16-rounded tensors -- the fraction of samples whose corner indices
    differ is asserted small, and every output pixel whose samples did not flip matches the oracle."""
    z, prm, m = _golden.load("n3s2_far")
    N, s, H, W = m["N"], m["s"], m["H"], m["W"]
    xb = _rnd(torch.from_numpy(z["x"]))
    p = oracle.LDConvParams(**{**prm.__dict__})
    for k in ("p_conv_weight", "p_conv_bias", "conv_weight", "bn_weight", "bn_bias", "running_mean", "running_var"):
        setattr(p, k, _rnd(torch.from_numpy(getattr(prm, k))))
    f = oracle.forward(xb, p, training=False, update_running=False)
    mod = E.LDConv(m["inc"], m["outc"], N, s)
    mod.load_state_dict({k: torch.from_numpy(np.array(z["param_" + k.replace(".", "_")])) for k in mod.state_dict()}, strict=True)
    mod.conv[1].eps = prm.eps
    dmod = mod.to(DEV).bfloat16().eval()
    xd = torch.from_numpy(xb).to(DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    # device offsets: the offset conv the module itself runs for this dtype / shape
    L = _lib.load()
    pr = dmod._prepared(torch.bfloat16, False)
    B = xb.shape[0]
    h, w = m["h"], m["w"]
    off = torch.empty((B, h, w, 2 * N), device=DEV, dtype=torch.float32)
    xh = xd.permute(0, 2, 3, 1).contiguous()
    _lib.check(L.ldconv_offset_conv_fwd(_ptr(xh), _ptr(pr.w_off), _ptr(pr.b_off), _ptr(off), B, m["inc"], H, W, N, s, _lib.BF16,
                                        _stream()), "ldconv_offset_conv_fwd")
    off_dev = off.permute(0, 3, 1, 2).contiguous().cpu().numpy()
    assert np.abs(off_dev - f["offset"]).max() <= 1e-3 * max(1.0, np.abs(f["offset"]).max())
    i_dev, _, _ = oracle.grid(off_dev, H, W, N, s)
    i_ref, _, _ = oracle.grid(f["offset"], H, W, N, s)
    flipped = (i_dev != i_ref).any(axis=-1)                   # (B,h,w,N)
    frac = float(flipped.mean())
    assert frac <= 0.01, frac
    with torch.no_grad():
        y = dmod(xd).float().cpu().numpy()
    keep = ~flipped.any(axis=-1)                              # (B,h,w): pixels none of whose samples flipped
    yk = y.transpose(0, 2, 3, 1)[keep]
    rk = f["out"].transpose(0, 2, 3, 1)[keep]
    assert keep.mean() >= 0.97
    assert _rel(yk, rk) <= 1e-2
