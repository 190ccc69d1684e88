"""GPU: the DEAL-YOLO-LD benchmark graph with the CUDA LDConv against the golden output of the reference DetectionModel
(tests/golden/model_deal_yolo_ld.npz, minted by oracle/gen_model_golden.py)."""
import os

import numpy as np
import pytest
import torch

from experiment_yolo_b200 import dealyolo
from experiment_yolo_b200.ldconv import LDConv as E_LDConv
from tests import _golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _load():
    z = np.load(os.path.join(_golden.GOLDEN_DIR, "model_deal_yolo_ld.npz"))
    model = dealyolo.DealYolo(nc=6)
    model.load_state_dict(dealyolo.seeded_state(model, seed=0), strict=True)
    return z, model


def test_full_model_fp32_matches_reference_output():
    """BASELINE config 1 (shrunk spatially): new module on the GPU vs the reference on the CPU, fp32, TF32 off."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    z, model = _load()
    model = model.to(DEV).eval()
    with torch.no_grad():
        y, feats = model(torch.from_numpy(z["x"]).to(DEV))
    y = y.cpu().numpy()
    ref = z["y"]
    # boxes are in pixels (up to ~128), class scores in [0,1]
    assert np.abs(y[:, 4:] - ref[:, 4:]).max() <= 1e-4
    assert np.abs(y[:, :4] - ref[:, :4]).max() <= 2e-3
    assert np.abs(feats[0].cpu().numpy() - z["feat0"]).max() <= 1e-3


def test_full_model_bf16_channels_last_close_to_reference_output():
    z, model = _load()
    model = dealyolo.channels_last_(model.to(DEV).bfloat16().eval())
    x = torch.from_numpy(z["x"]).to(DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y, _ = model(x)
    y = y.float().cpu().numpy()
    ref = z["y"]
    rel = np.linalg.norm(y - ref) / np.linalg.norm(ref)
    assert rel <= 5e-2, rel      # 27 bf16 layers deep; the LDConv layers alone are held to 1e-2 in test_gpu_parity.py


def test_full_model_train_mode_backward_runs():
    z, model = _load()
    model = model.to(DEV).train()
    x = torch.rand(2, 3, 64, 64, device=DEV)
    outs = model(x)
    loss = sum(o.float().square().mean() for o in outs)
    loss.backward()
    g = [p.grad for n, p in model.named_parameters() if "p_conv" in n or ".conv.0." in n]
    assert all(t is not None and bool(torch.isfinite(t).all()) for t in g)


def test_fused_engine_matches_reference_output_and_eager_graph():
    """engine.FusedDealYolo (Conv/C2f/SPPF/ScalSeq/Detect through the library's kernels, concat-free) against the golden output
    of the reference DetectionModel and against the eager bf16 graph."""
    from experiment_yolo_b200 import engine
    z, model = _load()
    model = dealyolo.channels_last_(model.to(DEV).bfloat16().eval())
    x = torch.from_numpy(z["x"]).to(DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    eng = engine.FusedDealYolo(model)
    y, _ = eng(x)
    torch.cuda.synchronize()
    with torch.no_grad():
        y_eager, _ = model(x)
    y, y_eager, ref = y.float().cpu().numpy(), y_eager.float().cpu().numpy(), z["y"]
    assert y.shape == ref.shape
    assert np.isfinite(y).all()
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) <= 5e-2
    assert np.linalg.norm(y - y_eager) / np.linalg.norm(y_eager) <= 5e-2
    # class scores are probabilities: compare them separately from the pixel-valued boxes
    assert np.abs(y[:, 4:] - ref[:, 4:]).max() <= 0.05


def test_fused_engine_batch_and_graph_capture():
    from experiment_yolo_b200 import engine
    z, model = _load()
    model = dealyolo.channels_last_(model.to(DEV).bfloat16().eval())
    eng = engine.FusedDealYolo(model)
    x = torch.rand(4, 3, 128, 160, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    y0, _ = eng(x)
    y0 = y0.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng(x)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y1, _ = eng(x)
    g.replay()
    torch.cuda.synchronize()
    assert tuple(y0.shape) == (4, 10, 32 * 40 + 16 * 20 + 8 * 10)
    assert torch.equal(y0, y1)


def test_pipelined_predictor_host_buffers_round_trip():
    """engine.PipelinedPredictor: uint8 host batches in, detections in host memory out, in submission order, equal to the
    synchronous executor on the same images."""
    from experiment_yolo_b200 import engine
    z, model = _load()
    model = dealyolo.channels_last_(model.to(DEV).bfloat16().eval())
    pred = engine.PipelinedPredictor(model, batch=2, imgsz=96)
    g = torch.Generator().manual_seed(9)
    batches = [torch.randint(0, 256, (2, 3, 96, 96), dtype=torch.uint8, generator=g).pin_memory() for _ in range(5)]
    outs = []
    for i, b in enumerate(batches):
        pred.submit(b)
        if i >= 1:
            outs.append(pred.result().clone())
    outs.append(pred.result().clone())
    eng = engine.FusedDealYolo(model)
    for b, o in zip(batches, outs):
        y, _ = eng(b.to(DEV))
        assert torch.equal(y.cpu(), o)
    # the uint8 path equals the float path on the normalised images
    y_f, _ = eng((batches[0].to(DEV).float() / 255.0).bfloat16())
    assert float((y_f.float().cpu() - outs[0].float()).abs().max()) <= 0.5


def test_fused_engine_accepts_a_model_built_with_foreign_ldconv_rows():
    """the executor recognises rows by class name / structure and converts reference-style LDConv rows in
    place (ldconv.convert): a graph whose LDConv rows are instances of ANOTHER class named `LDConv` with the reference's
    children (here: the eager CPU port under that name) gives the same output as the native graph."""
    from experiment_yolo_b200 import engine
    from oracle.ldconv_torch_port import LDConvTorchPort
    Foreign = type("LDConv", (LDConvTorchPort,), {})          # class name `LDConv`, not experiment_yolo_b200's type
    z, model = _load()
    foreign = dealyolo.DealYolo(nc=6, ldconv_cls=Foreign)
    foreign.load_state_dict(model.state_dict(), strict=True)
    assert not any(isinstance(m, E_LDConv) for m in foreign.modules())
    foreign = foreign.to(DEV).bfloat16().eval()
    model = dealyolo.channels_last_(model.to(DEV).bfloat16().eval())
    x = torch.from_numpy(z["x"]).to(DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    y_native, _ = engine.FusedDealYolo(model)(x)
    eng = engine.FusedDealYolo(foreign)
    assert sum(isinstance(m, E_LDConv) for m in foreign.modules()) == 10
    y_foreign, _ = eng(x)
    torch.cuda.synchronize()
    assert torch.equal(y_native, y_foreign)


def test_engine_add_row_sums_all_inputs_through_the_library():
    from experiment_yolo_b200 import engine
    g = torch.Generator(device=DEV).manual_seed(3)
    wide = torch.randn((2, 12, 20, 64), device=DEV, generator=g).bfloat16()
    xs = [torch.randn((2, 12, 20, 32), device=DEV, generator=g).bfloat16(), wide[..., 16:48],
          torch.randn((2, 12, 20, 32), device=DEV, generator=g).bfloat16()]
    for n in (2, 3):
        y = engine.add_nhwc(xs[:n])
        want = torch.stack([t.float() for t in xs[:n]]).sum(0).bfloat16()
        assert torch.equal(y, want)
    five = xs + xs[:2]
    y = engine.add_nhwc(five)
    assert float((y.float() - torch.stack([t.float() for t in five]).sum(0)).abs().max()) <= 0.07


def test_detect_decode_in_the_gemm_epilogue_is_bit_identical_to_conv_then_decode():
    """engine._Detect.decode_in_epilogue: the last 1x1 conv of every Detect branch decodes its own accumulator rows
    (ldconv_conv1x1_detect_fwd) -- same logits rounding, same DFL / dist2bbox / sigmoid arithmetic as ldconv_conv1x1_bn_act_fwd +
    ldconv_detect_decode (head.py:55-77), so the decoded output is bit-identical; non-square maps, batch 3."""
    from experiment_yolo_b200 import _lib, engine
    z, model = _load()
    model = dealyolo.channels_last_(model.to(DEV).bfloat16().eval())
    x = torch.rand(3, 3, 96, 160, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    eng = engine.FusedDealYolo(model)
    try:
        engine._Detect.decode_in_epilogue = False
        _lib.call_counts.clear()
        y_ref, feats = eng(x)
        assert "ldconv_detect_decode" in _lib.call_counts and feats is not None
        engine._Detect.decode_in_epilogue = True
        _lib.call_counts.clear()
        y, none = eng(x)
        assert "ldconv_detect_decode" not in _lib.call_counts and _lib.call_counts["ldconv_conv1x1_detect_fwd"] == 6 and none is None
    finally:
        engine._Detect.decode_in_epilogue = True
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref)


@pytest.mark.parametrize("B,C,H,W", [(4, 32, 24, 40), (2, 16, 33, 17), (3, 96, 10, 10), (2, 64, 160, 160)])
def test_fused_bn_silu_training_op_vs_torch(B, C, H, W):
    """ldconv.bn_silu_train (col_stats -> bn_finalize -> bn_act_apply; backward reduce + apply) against torch's BatchNorm2d + SiLU in
    fp32 on the same bf16 pre-activation: output, running statistics, and the gradients of the input / gamma / beta."""
    from experiment_yolo_b200.ldconv import bn_silu_train
    torch.manual_seed(C + H)
    pre = (torch.randn(B, C, H, W, device=DEV) * 1.5 + 0.3).bfloat16().contiguous(memory_format=torch.channels_last)
    gout = torch.randn(B, C, H, W, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(C, eps=1e-3, momentum=0.03).to(DEV).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    ref_bn = torch.nn.BatchNorm2d(C, eps=1e-3, momentum=0.03).to(DEV).train()
    ref_bn.load_state_dict(bn.state_dict())
    x1 = pre.clone().requires_grad_(True)
    y1 = bn_silu_train(x1, bn)
    assert y1 is not None and y1.dtype == torch.bfloat16 and y1.is_contiguous(memory_format=torch.channels_last)
    y1.backward(gout)
    x2 = pre.float().clone().requires_grad_(True)
    y2 = torch.nn.functional.silu(ref_bn(x2))
    y2.backward(gout.float())
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
    assert rel(y1, y2) <= 4e-3
    assert rel(x1.grad, x2.grad) <= 1e-2
    assert rel(bn.weight.grad, ref_bn.weight.grad) <= 2e-3 and rel(bn.bias.grad, ref_bn.bias.grad) <= 2e-3
    assert torch.allclose(bn.running_mean, ref_bn.running_mean, atol=1e-5) and torch.allclose(bn.running_var, ref_bn.running_var, rtol=1e-4, atol=1e-6)
    assert int(bn.num_batches_tracked) == 1


def test_fused_engine_on_a_model_whose_conv_blocks_were_fused_like_model_fuse():
    """The reference's `model.fuse()` (nn/tasks.py:168-195) folds BatchNorm into every `Conv` block's convolution and deletes `bn`;
    the executor must take such blocks too (Conv2d with bias + SiLU).  Folding here is done in fp32 exactly like
    `fuse_conv_and_bn` (utils/torch_utils.py), so the executor's own fold of the unfused model agrees to bf16 rounding of the weights."""
    from experiment_yolo_b200 import engine
    z, model = _load()
    fused = copy_model(model)
    for m in fused.modules():
        if type(m).__name__ == "Conv" and hasattr(m, "bn"):
            conv, bn = m.conv, m.bn
            w = conv.weight.detach().clone()
            scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
            new = torch.nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, bias=True)
            new.weight.data = w * scale.view(-1, 1, 1, 1)
            new.bias.data = bn.bias.detach() - bn.running_mean.detach() * scale
            m.conv = new
            del m.bn
    model = dealyolo.channels_last_(model.to(DEV).bfloat16().eval())
    fused = fused.to(DEV).eval()      # fp32 parameters: the executor rounds the folded weights itself
    x = torch.from_numpy(z["x"]).to(DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    y0, _ = engine.FusedDealYolo(model)(x)
    y1, _ = engine.FusedDealYolo(fused)(x)
    torch.cuda.synchronize()
    a, b = y1.float().cpu().numpy(), y0.float().cpu().numpy()
    assert np.isfinite(a).all()
    assert np.linalg.norm(a - b) / np.linalg.norm(b) <= 2e-2
    assert np.linalg.norm(a - z["y"]) / np.linalg.norm(z["y"]) <= 5e-2


def copy_model(model):
    import copy
    return copy.deepcopy(model)


@pytest.mark.parametrize("B,C,H,W,f,sliced", [(2, 32, 20, 12, 2, False), (1, 64, 7, 9, 2, True), (2, 16, 5, 6, 4, True), (1, 8, 3, 3, 1, False)])
def test_library_upsample_nearest_forward_backward_vs_torch(B, C, H, W, f, sliced):
    """dealyolo.Upsample (the YAML's nn.Upsample [None, f, nearest] rows) on CUDA bf16: forward bit-equal to torch's nearest
    interpolation, backward = the f x f block sums of the gradient (fp32 sum, one bf16 rounding), also when the gradient arrives as
    a channel slice of a wider channels_last tensor (what the Concat behind the row hands back)."""
    from experiment_yolo_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(B * 100 + C + f)
    x = torch.randn((B, C, H, W), device=DEV, generator=g).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    up = dealyolo.Upsample(None, f, "nearest")
    _lib.call_counts.clear()
    y = up(x)
    assert "ldconv_upsample_nearest" in _lib.call_counts
    ref = torch.nn.functional.interpolate(x.detach().float(), scale_factor=f, mode="nearest")
    assert torch.equal(y.float(), ref)
    if sliced:
        wide = torch.randn((B, C + 24, H * f, W * f), device=DEV, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
        gout = wide[:, 8:8 + C]
    else:
        gout = torch.randn((B, C, H * f, W * f), device=DEV, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    y.backward(gout)
    xr = x.detach().float().requires_grad_(True)
    torch.nn.functional.interpolate(xr, scale_factor=f, mode="nearest").backward(gout.float())
    assert "ldconv_upsample_nearest_bwd" in _lib.call_counts
    assert torch.equal(x.grad.float(), xr.grad.bfloat16().float())


def test_library_add_row_with_autograd_vs_torch():
    """dealyolo.Add on CUDA bf16 maps (one dense, one a channel slice): the library sum equals torch's fp32 sum rounded once, and the
    gradient reaches every input unchanged."""
    from experiment_yolo_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(5)
    a = torch.randn((2, 32, 12, 20), device=DEV, generator=g).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wide = torch.randn((2, 48, 12, 20), device=DEV, generator=g).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    c = torch.randn((2, 32, 12, 20), device=DEV, generator=g).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    _lib.call_counts.clear()
    y = dealyolo.Add()([a, wide[:, 16:48], c])
    assert "ldconv_add_nhwc" in _lib.call_counts
    ref = (a.detach().float() + wide.detach()[:, 16:48].float() + c.detach().float()).bfloat16()
    assert torch.equal(y, ref)
    gout = torch.randn_like(y)
    y.backward(gout)
    assert torch.equal(a.grad, gout) and torch.equal(c.grad, gout)
    assert torch.equal(wide.grad[:, 16:48], gout) and float(wide.grad[:, :16].abs().max()) == 0.0


def test_scalseq_training_tail_through_the_library_vs_torch():
    """dealyolo.ScalSeq in training mode under bf16 autocast: the tail after the three 1x1 Convs (up-sampling, stack, Conv3d 1x1x1,
    BatchNorm3d with batch statistics, LeakyReLU(0.1), MaxPool3d over the depth axis; nn/extra_modules/block.py:3432-3443) through
    train_ops.scalseq_tail against torch's own ops on the same module: output, input / parameter gradients and running statistics."""
    import copy
    from experiment_yolo_b200 import _lib
    torch.manual_seed(11)
    ref = dealyolo.ScalSeq([32, 64, 128], 32).to(DEV).train()
    with torch.no_grad():
        ref.bn.weight.uniform_(0.5, 1.5)
        ref.bn.bias.normal_(0, 0.2)
        ref.conv3d.bias.normal_(0, 0.5)
    lib = copy.deepcopy(ref)
    xs = [torch.randn(2, c, h, w, device=DEV).contiguous(memory_format=torch.channels_last) for c, h, w in ((32, 16, 24), (64, 8, 12), (128, 4, 6))]
    outs, grads = {}, {}
    for name, mod, fused in (("ref", ref, False), ("lib", lib, True)):
        dealyolo.ScalSeq.fused_train_tail = fused
        try:
            ins = [t.clone().requires_grad_(True) for t in xs]
            _lib.call_counts.clear()
            with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                y = mod(ins)
            assert ("ldconv_ssff_max_fwd" in _lib.call_counts) == fused
            g = torch.randn(y.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
            y.float().backward(g)
            outs[name] = y.float().detach()
            grads[name] = [t.grad.float() for t in ins] + [mod.conv3d.weight.grad.float(), mod.bn.weight.grad.float(), mod.bn.bias.grad.float(),
                                                            mod.conv1.conv.weight.grad.float()]
        finally:
            dealyolo.ScalSeq.fused_train_tail = True
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-12))
    assert rel(outs["lib"], outs["ref"]) <= 1e-2
    for a, b in zip(grads["lib"], grads["ref"]):
        assert rel(a, b) <= 4e-2, (a.shape, rel(a, b))
    assert rel(lib.bn.running_mean, ref.bn.running_mean) <= 2e-3 and rel(lib.bn.running_var, ref.bn.running_var) <= 2e-3
    assert float(lib.conv3d.bias.grad.abs().max()) <= 1e-2 * max(1.0, float(ref.conv3d.bias.grad.abs().max()) * 100)
    assert int(lib.bn.num_batches_tracked) == int(ref.bn.num_batches_tracked) == 1
