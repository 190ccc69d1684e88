"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/ldconv_b200.h declares,
and the Python module keeps the reference's surface (conv.py:350-359: constructor, children, buffer, state_dict)."""
import copy
import ctypes
import os
import pickle
import re
import sys
import types

import numpy as np
import pytest
import torch

import experiment_yolo_b200 as E
from experiment_yolo_b200 import _lib, ldconv
from oracle import oracle
from tests import _golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ldconv_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ldconv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 17
    lib = ctypes.CDLL(_lib.LIB_PATH) if "torch" in sys.modules else None
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ldconv_b200.h but not exported by libldconv_b200.so"
    assert sorted(_lib.SIGNATURES) == names, "the ctypes table must mirror the header one to one"


def test_version_and_error_plumbing_without_gpu():
    L = _lib.load()
    assert L.ldconv_version() == 1
    # no compute: only argument validation, which happens before any CUDA call
    rc = L.ldconv_gather_fwd(None, None, None, None, None, None, 1, 8, 4, 4, 3, 1, 0, None)
    assert rc == -1 and b"null pointer" in L.ldconv_last_error()
    rc = L.ldconv_offset_conv_fwd(None, None, None, None, 1, 8, 4, 4, 99, 1, 0, None)
    assert rc == -1 and b"num_param" in L.ldconv_last_error()
    rc = L.ldconv_gemm_fwd(None, None, None, None, None, None, None, None, 4, 8, 8, 1, 7, None)
    assert rc == -1 and b"dtype" in L.ldconv_last_error()


@pytest.mark.parametrize("N", range(1, 17))
def test_p_n_helper_matches_oracle(N):
    want = oracle.p_n(N).reshape(-1).tolist()
    assert _lib.p_n_table(N) == want
    assert ldconv.base_grid(N).reshape(-1).tolist() == want
    assert ldconv.base_grid(N).dtype == torch.int64 and tuple(ldconv.base_grid(N).shape) == (1, 2 * N, 1, 1)


@pytest.mark.parametrize("name", _golden.case_names())
def test_state_dict_layout_matches_reference(name):
    """Keys, order, shapes and dtypes of state_dict() equal the reference module's."""
    z, prm, m = _golden.load(name)
    mod = E.LDConv(m["inc"], m["outc"], m["N"], m["s"])
    sd = mod.state_dict()
    ref_keys = ["p_n", "conv.0.weight", "conv.1.weight", "conv.1.bias", "conv.1.running_mean", "conv.1.running_var",
                "conv.1.num_batches_tracked", "p_conv.weight", "p_conv.bias"]
    assert list(sd.keys()) == ref_keys
    for k in ref_keys:
        g = z["param_" + k.replace(".", "_")]
        assert tuple(sd[k].shape) == tuple(g.shape), k
        assert str(sd[k].dtype).replace("torch.", "") == str(g.dtype), k
    assert torch.equal(sd["p_n"], torch.from_numpy(z["param_p_n"]))
    assert float(sd["p_conv.weight"].abs().max()) == 0.0          # conv.py:357
    assert mod.conv[0].bias is None                                # conv.py:351,355
    assert mod.num_param == m["N"] and mod.stride == m["s"]


def test_same_seed_same_initialisation_as_reference_order():
    """Construction order matches conv.py:355-357, so the RNG stream is consumed identically: conv weight, then p_conv
    (whose weight is then zeroed, bias kept)."""
    torch.manual_seed(7)
    a = E.LDConv(6, 10, 5, 2)
    torch.manual_seed(7)
    conv = torch.nn.Conv2d(6, 10, kernel_size=(5, 1), stride=(5, 1), bias=None)
    pconv = torch.nn.Conv2d(6, 10, kernel_size=3, padding=1, stride=2)
    assert torch.equal(a.conv[0].weight, conv.weight)
    assert torch.equal(a.p_conv.bias, pconv.bias)


def test_cpu_tensor_raises_the_error_the_reference_probe_expects():
    """nn/tasks.py:317-321 retries on the GPU when the RuntimeError text contains 'CUDA tensor'."""
    mod = E.LDConv(4, 8, 3, 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        mod(torch.zeros(2, 4, 8, 8))


def test_deepcopy_pickle_and_dtype_casts_keep_the_state():
    mod = E.LDConv(4, 8, 3, 1)
    mod2 = copy.deepcopy(mod)
    assert list(mod2.state_dict()) == list(mod.state_dict())
    mod3 = pickle.loads(pickle.dumps(mod))
    assert torch.equal(mod3.conv[0].weight, mod.conv[0].weight)
    half = copy.deepcopy(mod).bfloat16()
    assert half.p_n.dtype == torch.int64 and half.conv[0].weight.dtype == torch.bfloat16
    mod.load_state_dict(mod2.state_dict(), strict=True)


def test_yaml_hook_install_rebinds_the_four_reference_modules(monkeypatch):
    """parse_model resolves the YAML name through module globals (nn/tasks.py:813); install() rebinds `LDConv` in the four
    modules that hold it.  ultralytics itself is absent here and on the GPU box, so the package tree is faked."""
    class RefLDConv(torch.nn.Module):
        pass
    names = ["ultralytics", "ultralytics.nn", "ultralytics.nn.modules", "ultralytics.nn.modules.conv",
             "ultralytics.nn.modules.block", "ultralytics.nn.tasks"]
    for n in names:
        mod = types.ModuleType(n)
        mod.__path__ = []
        monkeypatch.setitem(sys.modules, n, mod)
    for n in ldconv._REF_MODULES:
        sys.modules[n].LDConv = RefLDConv
    patched = E.install()
    assert sorted(patched) == sorted(ldconv._REF_MODULES)
    for n in ldconv._REF_MODULES:
        assert sys.modules[n].LDConv is E.LDConv
    # the YAML row [-1, 1, LDConv, [c2, num_param, stride]] -> m(*[c1, c2, num_param, stride])  (nn/tasks.py:864,1046)
    layer = sys.modules["ultralytics.nn.tasks"].LDConv(*[16, 32, 3, 2])
    assert isinstance(layer, E.LDConv) and layer.p_conv.stride == (2, 2)


def test_convert_swaps_class_in_place():
    class LDConv(torch.nn.Module):          # stands in for the reference class (same name, same children)
        def __init__(self):
            super().__init__()
            src = E.LDConv(4, 8, 3, 2)
            self.num_param, self.stride = 3, 2
            self.conv, self.p_conv = src.conv, src.p_conv
            self.register_buffer("p_n", src.p_n.clone())
    holder = torch.nn.Sequential(LDConv())
    keys = list(holder.state_dict())
    ldconv.convert(holder)
    assert isinstance(holder[0], E.LDConv) and list(holder.state_dict()) == keys


def test_engine_pixel_packing_rule_and_block_diagonal_weights():
    """Host logic of the pixel-packed 1x1 conv (engine._pack_factor / _packed, ldconv_conv1x1_bn_act_packed_fwd): which layers are
    packed, and that P pixels times block_diag(W, ..., W) is the same product as the plain conv (the arithmetic the GPU test
    checks bit for bit)."""
    from experiment_yolo_b200 import engine
    pf = engine._pack_factor
    assert pf(32, 32, 1638400, 32, 48) == 4            # C2f.cv1 at P2: dense input, output slice of the 48-wide concat buffer
    assert pf(48, 32, 1000, 48, 32) == 4
    assert pf(64, 64, 1000, 64, 64) == 1               # 64-wide inputs stay on the plain kernel (staged stores)
    assert pf(32, 32, 1001, 32, 32) == 1               # rows not divisible by P
    assert pf(32, 32, 1002, 32, 32) == 2
    assert pf(32, 32, 1000, 64, 32) == 1               # input is a channel slice: P pixels are not one contiguous row
    assert pf(32, 48, 1000, 32, 48) == 1               # Cout not a power of two
    assert pf(32, 64, 1000, 32, 64) == 2               # P * Cout <= 128
    conv = torch.nn.Conv2d(16, 32, 1, bias=False)
    bn = torch.nn.BatchNorm2d(32).eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.2)
        bn.running_var.uniform_(0.5, 1.5)
    f = engine._Folded(conv, bn)
    wp, sc, sh = engine._packed(f, 4)
    assert engine._packed(f, 4)[0] is wp                # cached on the _Folded
    assert tuple(wp.shape) == (128, 64) and tuple(sc.shape) == (128,) and torch.equal(sc[:32], sc[96:]) and torch.equal(sh[:32], sh[32:64])
    x = torch.randn(8, 16).bfloat16().float()
    plain = (x @ f.w.float().t()) * f.scale + f.shift                                              # (8 pixels, 32)
    packed = (x.reshape(2, 64) @ wp.float().t()) * sc + sh                                          # (2 rows, 4 x 32)
    # (the CPU matmul blocks K = 64 and K = 16 differently, so this host-side statement is equal up to fp32 rounding)
    assert torch.allclose(packed.reshape(8, 32), plain, rtol=0, atol=1e-5)


_REAL_INSTALL_SCRIPT = r'''
import json, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, {root!r})
import torch
from oracle.gen_model_golden import load_reference_tasks, REF_YAML
T = load_reference_tasks()                                  # the unmodified reference, third-party roots stubbed
import experiment_yolo_b200 as E
from experiment_yolo_b200 import ldconv
ref_cls = T.LDConv
torch.manual_seed(0)
ref_model, _ = T.parse_model(T.yaml_model_load(REF_YAML) | {{"nc": 6}}, ch=3, verbose=False)
patched = E.install()
torch.manual_seed(0)
new_model, save = T.parse_model(T.yaml_model_load(REF_YAML) | {{"nc": 6}}, ch=3, verbose=False)
rows = [type(m).__module__ + "." + type(m).__name__ for m in new_model if type(m).__name__ == "LDConv"]
ref_sd, new_sd = ref_model.state_dict(), new_model.state_dict()
same_keys = list(ref_sd) == list(new_sd)
same_vals = same_keys and all(torch.equal(ref_sd[k], new_sd[k]) for k in ref_sd)
probe = ""
try:
    new_model[0](torch.zeros(2, 3, 64, 64))
except RuntimeError as e:
    probe = str(e)
from experiment_yolo_b200 import engine
det = T.DetectionModel.__new__(T.DetectionModel)           # the executor's row recognition on the REAL reference modules
torch.nn.Module.__init__(det)
det.model = ref_model
kinds = engine.FusedDealYolo.recognise(det)
import ultralytics.nn.modules.block as RB
bott = RB.Bottleneck_LDConv(16, 16)                        # block.py:628-629 picks the rebound name up too
print(json.dumps({{"patched": patched, "rows": rows, "same_keys": same_keys, "same_vals": same_vals, "probe": probe,
                  "params": sum(p.numel() for p in new_model.parameters()),
                  "ref_is_other_class": ref_cls is not E.LDConv, "kinds": kinds,
                  "bottleneck": [type(bott.cv1).__module__, type(bott.cv2).__module__]}}))
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference/ultralytics"), reason="the reference tree only exists in the authoring container")
def test_install_against_the_real_reference_parse_model(tmp_path):
    """`install()` on the REAL `ultralytics.nn.tasks` (stub importer of SURVEY.md App. D for the absent
    third-party roots), in a subprocess so the reference import cannot leak into the other tests: `parse_model`
    (nn/tasks.py:813-864) builds all ten LDConv rows of yolov8-LD-P2.yaml with this class, the state_dict equals the
    reference-built model's under the same seed (keys, order, values), and a CPU tensor raises the 'CUDA tensor' error the
    stride probe retries on (nn/tasks.py:317-321)."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, YOLO_CONFIG_DIR=str(tmp_path), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", _REAL_INSTALL_SCRIPT.format(root=root)], capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert sorted(out["patched"]) == sorted(ldconv._REF_MODULES)
    assert out["ref_is_other_class"]
    assert out["rows"] == ["experiment_yolo_b200.ldconv.LDConv"] * 10
    assert out["same_keys"] and out["same_vals"]
    assert out["params"] == 918304
    assert "CUDA tensor" in out["probe"]
    assert out["bottleneck"] == ["experiment_yolo_b200.ldconv"] * 2
    # engine.FusedDealYolo recognises every row of the reference-built graph (yolov8-LD-P2.yaml:14-52)
    assert out["kinds"] == ["ldconv", "ldconv", "c2f", "ldconv", "c2f", "ldconv", "c2f", "sppf", "ldconv", "up", "ldconv", "cat",
                            "c2f", "ldconv", "up", "ldconv", "cat", "c2f", "ldconv", "cat", "c2f", "ldconv", "cat", "c2f",
                            "scalseq", "add", "detect"]


def test_train_ops_host_logic_on_cpu():
    """train_ops: the NHWC view helper takes dense channels_last tensors and channel slices without a copy and copies anything else;
    the library-backed training ops decline CPU tensors (None -> the caller runs torch's op), so the CPU graph is unchanged."""
    import torch
    from experiment_yolo_b200 import dealyolo, train_ops
    x = torch.randn(2, 32, 6, 5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    v, ld = train_ops.nhwc_view(x)
    assert ld == 32 and v.data_ptr() == x.data_ptr() and v.shape == (2, 6, 5, 32)
    wide = torch.randn(2, 48, 6, 5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    v, ld = train_ops.nhwc_view(wide[:, 8:40])
    assert ld == 48 and v.data_ptr() == wide.data_ptr() + 8 * 2
    v, ld = train_ops.nhwc_view(torch.randn(2, 32, 6, 5).to(torch.bfloat16))          # NCHW-contiguous: copied
    assert ld == 32 and v.is_contiguous()
    v, ld = train_ops.nhwc_view(wide[:, 3:35])                                        # slice start not 16-byte aligned: copied
    assert ld == 32 and v.is_contiguous()
    assert train_ops.upsample_nearest(x, 2) is None and train_ops.add_maps([x, x]) is None
    up = dealyolo.Upsample(None, 2, "nearest")
    assert torch.equal(up(x.float()), torch.nn.functional.interpolate(x.float(), scale_factor=2, mode="nearest"))
    assert torch.equal(dealyolo.Add()([x.float(), x.float()]), 2 * x.float())
    seq = dealyolo.ScalSeq([32, 64, 128], 32).train()
    y = seq([torch.randn(2, 32, 8, 8), torch.randn(2, 64, 4, 4), torch.randn(2, 128, 2, 2)])
    assert y.shape == (2, 32, 8, 8)
