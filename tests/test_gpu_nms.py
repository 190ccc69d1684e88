"""GPU: the device-side non_max_suppression (csrc/ldconv_nms.cu through experiment_yolo_b200.nms) against the golden vectors
minted from the reference function (ultralytics/utils/ops.py:292-427 + soft_nms :260-290; oracle/gen_nms_golden.py) and
against the numpy oracle on larger seeded inputs."""
import os

import numpy as np
import pytest
import torch

from experiment_yolo_b200 import nms
from oracle import nms_oracle
from oracle.gen_nms_golden import synth
from tests.test_nms_cpu import CASES, load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _compare(got, want):
    assert [tuple(g.shape) for g in got] == [tuple(w.shape) for w in want]
    for g, w in zip(got, want):
        g = g.cpu().numpy()
        assert np.array_equal(g[:, 5], w[:, 5])
        assert np.array_equal(g[:, :4], w[:, :4])
        assert np.abs(g[:, 4] - w[:, 4]).max(initial=0.0) <= 3e-7


@pytest.mark.parametrize("name", CASES)
def test_nms_matches_reference_fixture(name):
    pred, kw, want = load(name)
    got = nms.non_max_suppression(torch.from_numpy(pred).to(DEV), **kw)
    _compare(got, want)


def test_nms_bf16_input_equals_fp32_input_of_the_same_values():
    pred, kw, want = load("bf16_agnostic")            # values are bf16-representable
    got = nms.non_max_suppression(torch.from_numpy(pred).to(DEV).bfloat16(), **kw)
    _compare(got, want)


@pytest.mark.parametrize("B,A,objects,hot", [(8, 33600, 200, 0.02), (2, 33600, 40, 0.1), (64, 8400, 30, 0.03)])
def test_nms_head_sized_inputs_vs_oracle(B, A, objects, hot):
    """the model's real anchor count (33600 at 640x640), hundreds to thousands of candidates per image"""
    pred = synth(100 + B, B, A, 6, objects, hot, dtype=torch.bfloat16)
    kw = dict(conf_thres=0.25, iou_thres=0.45)
    got = nms.non_max_suppression(pred.to(DEV).bfloat16(), **kw)
    want = nms_oracle.non_max_suppression(pred.numpy()[: min(B, 4)], **kw)      # the oracle is a Python loop: first images only
    _compare(got[: len(want)], want)
    assert all(0 < g.shape[0] <= 300 for g in got)


def test_nms_rejects_what_it_does_not_cover():
    pred = torch.rand(1, 10, 64, device=DEV)
    with pytest.raises(NotImplementedError):
        nms.non_max_suppression(pred, classes=[1])
    with pytest.raises(NotImplementedError):
        nms.non_max_suppression(pred, multi_label=True)
    with pytest.raises(RuntimeError):
        nms.non_max_suppression(pred.cpu())
    with pytest.raises(RuntimeError, match="max_nms"):
        nms.non_max_suppression(torch.rand(1, 10, 64, device=DEV) * 0.5 + 0.5, max_nms=10)


def test_pipelined_predictor_with_device_nms():
    """engine.PipelinedPredictor(nms=...): detections (B, max_det, 6) + counts come back instead of the 43 MB head output
    and equal the eager API on the synchronous executor's output."""
    from experiment_yolo_b200 import dealyolo, engine
    model = dealyolo.DealYolo(nc=6)
    model.load_state_dict(dealyolo.seeded_state(model, seed=0), strict=True)
    model = dealyolo.channels_last_(model.to(DEV).bfloat16().eval())
    eng = engine.FusedDealYolo(model)
    g = torch.Generator().manual_seed(11)
    batches = [torch.randint(0, 256, (2, 3, 96, 96), dtype=torch.uint8, generator=g).pin_memory() for _ in range(3)]
    y0, _ = eng(batches[0].to(DEV))
    thr = float(y0[:, 4:].float().amax(1).flatten().quantile(0.9))      # random weights: take the top 10 % as candidates
    kw = dict(conf_thres=thr, iou_thres=0.45, max_det=50)
    pred = engine.PipelinedPredictor(model, batch=2, imgsz=96, nms=kw)
    assert pred.d2h_bytes == 2 * 50 * 6 * 4 + 2 * 4
    outs = []
    for i, b in enumerate(batches):
        pred.submit(b)
        if i >= 1:
            outs.append([t.clone() for t in pred.detections(pred.result())])
    outs.append([t.clone() for t in pred.detections(pred.result())])
    for b, o in zip(batches, outs):
        y, _ = eng(b.to(DEV))
        want = nms.non_max_suppression(y, **kw)
        assert len(o) == len(want) == 2
        for a, w in zip(o, want):
            assert torch.equal(a, w.cpu())
        assert sum(t.shape[0] for t in o) > 0
