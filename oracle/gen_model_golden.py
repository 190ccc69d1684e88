"""Mint the full-model golden vector FROM THE REFERENCE's own DetectionModel (authoring container only).

Imports the unmodified reference (`/root/reference/ultralytics`) with the stub importer of SURVEY.md Appendix D (ten
third-party roots that are absent here are replaced by MagicMock modules), builds
`DetectionModel('cfg/models/yolov8-LD-P2.yaml', ch=3, nc=6)` (nn/tasks.py:275-333), loads the deterministic synthetic
weights of experiment_yolo_b200.dealyolo.seeded_state (reproducible without the reference), runs an eval forward in
fp32 on the CPU on a seeded non-square image and stores the decoded head output under tests/golden/.
It also cross-checks, right here, that the benchmark graph (experiment_yolo_b200.dealyolo.DealYolo) and the eager CPU
port of LDConv (oracle/ldconv_torch_port.py) reproduce the reference: identical state_dict keys / shapes, identical
output, and comparable wall-clock time (the port is the `cpu_baseline` bench.py times on the GPU box).

    python oracle/gen_model_golden.py

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import json
import os
import sys
import tempfile
import time
import warnings
from unittest.mock import MagicMock

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "model_deal_yolo_ld.npz")
OUT640 = os.path.join(ROOT, "tests", "golden", "model_deal_yolo_ld_640.npz")
REF_YAML = "/root/reference/ultralytics/cfg/models/yolov8-LD-P2.yaml"
STUB_ROOTS = ("matplotlib", "timm", "thop", "mmcv", "mmengine", "efficientnet_pytorch", "mamba_ssm", "DCNv3", "DCNv4",
              "swattention")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__name__, m.__path__, m.__spec__, m.__loader__ = spec.name, [], spec, self
        return m

    def exec_module(self, module):
        pass


def load_reference_tasks():
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp())
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    sys.meta_path.append(_StubFinder())
    import ultralytics.nn.tasks as T
    return T


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(8)
    T = load_reference_tasks()
    from experiment_yolo_b200 import dealyolo
    from oracle.ldconv_torch_port import LDConvTorchPort

    ref = T.DetectionModel(REF_YAML, ch=3, nc=6, verbose=False)
    mine = dealyolo.DealYolo(nc=6, ldconv_cls=LDConvTorchPort)
    assert [float(v) for v in ref.stride] == [float(v) for v in mine.stride], (ref.stride, mine.stride)
    rk, mk = list(ref.state_dict().keys()), list(mine.state_dict().keys())
    assert rk == mk, [k for k in rk if k not in mk][:5] + [k for k in mk if k not in rk][:5]
    for k in rk:
        assert ref.state_dict()[k].shape == mine.state_dict()[k].shape, k
    n_params = sum(p.numel() for p in ref.parameters())
    sd = dealyolo.seeded_state(mine, seed=0)
    ref.load_state_dict(sd, strict=True)
    mine.load_state_dict(sd, strict=True)
    ref.eval(); mine.eval()

    g = torch.Generator().manual_seed(123)
    x = torch.rand(1, 3, 96, 128, generator=g)
    with torch.no_grad():
        y_ref, feats_ref = ref(x)
        y_mine, feats_mine = mine(x)
    diff = float((y_ref - y_mine).abs().max())
    print(f"params {n_params}  out {tuple(y_ref.shape)}  |ref - harness(port)| max = {diff:.3e}")
    assert diff == 0.0, "the benchmark graph + eager port must reproduce the reference bit for bit on the CPU"

    # BASELINE config 1 at its real size: torch.rand(1,3,640,640, seed 0), fp32, eval.  The input is regenerated
    # from the seed by the tests (a checksum and a strided sample are stored to catch a generator change); only the decoded
    # head output and the three LDConv-heavy feature statistics are stored.
    g0 = torch.Generator().manual_seed(0)
    x640 = torch.rand(1, 3, 640, 640, generator=g0)
    with torch.no_grad():
        y640, f640 = ref(x640)
        y640_mine, _ = mine(x640)
    assert float((y640 - y640_mine).abs().max()) == 0.0
    np.savez_compressed(OUT640, y=y640.numpy(), x_sum=np.float64(x640.double().sum().item()),
                        x_probe=x640.reshape(-1)[::4099].numpy(), feat_absmean=np.array([float(f.abs().mean()) for f in f640]))
    print("wrote", OUT640, os.path.getsize(OUT640) // 1024, "KiB")

    # train-mode outputs (raw head maps, batch statistics) for the training-step harness
    ref.train(); mine.train()
    x2 = torch.rand(2, 3, 64, 64, generator=g)
    t_ref = ref(x2)
    t_mine = mine(x2)
    for a, b in zip(t_ref, t_mine):
        assert float((a - b).abs().max()) <= 1e-5
    ref.eval(); mine.eval()

    # wall-clock: reference vs port on the headline shape (B=1, 640x640), best of 3
    xb = torch.rand(1, 3, 640, 640, generator=g)
    times = {}
    with torch.inference_mode():
        for name, m in (("reference", ref), ("port", mine)):
            m(xb)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); m(xb); best = min(best, time.perf_counter() - t0)
            times[name] = best
    print(f"CPU forward 1x3x640x640, {torch.get_num_threads()} threads: reference {times['reference']*1e3:.1f} ms, "
          f"port {times['port']*1e3:.1f} ms")

    np.savez_compressed(OUT, x=x.numpy(), y=y_ref.numpy(), feat0=feats_ref[0].numpy(), n_params=np.int64(n_params),
                        strides=np.array([float(v) for v in ref.stride]),
                        cpu_ms_reference=np.float64(times["reference"] * 1e3), cpu_ms_port=np.float64(times["port"] * 1e3))
    with open(os.path.join(ROOT, "tests", "golden", "MODEL_MANIFEST.json"), "w") as f:
        json.dump({"generator": "oracle/gen_model_golden.py", "reference": REF_YAML, "params": int(n_params),
                   "weights": "experiment_yolo_b200.dealyolo.seeded_state(seed=0)", "input": "torch.rand(1,3,96,128), seed 123",
                   "cpu_ms_640_reference": times["reference"] * 1e3, "cpu_ms_640_port": times["port"] * 1e3,
                   "threads": torch.get_num_threads(), "torch": torch.__version__}, f, indent=1)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
