"""Times ldconv_conv3x3_bn_act_fwd at the 3x3 conv shapes of the DEAL-YOLO-LD step (batch 64): CUDA events, L2 flushed, median."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import _lib  # noqa: E402

SHAPES = [(16, 16, 160), (32, 32, 80), (64, 64, 40), (32, 96, 160), (64, 64, 160), (32, 32, 160), (64, 96, 80), (64, 64, 80), (128, 64, 40), (32, 16, 160)]


def main():
    L = _lib.load()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    B = 64
    tot = 0.0
    for ci, co, hw in SHAPES:
        x = torch.randn((B, hw, hw, ci), device=dev).bfloat16()
        w = (torch.randn((co, 9 * ci), device=dev) * 0.05).bfloat16()
        sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
        out = torch.empty((B, hw, hw, co), device=dev, dtype=torch.bfloat16)
        ms = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            _lib.check(L.ldconv_conv3x3_bn_act_fwd(x.data_ptr(), ci, w.data_ptr(), sc.data_ptr(), sh.data_ptr(), None, 0, out.data_ptr(), co,
                                                   B, ci, hw, hw, co, 1, _lib.ACT_SILU, _lib.BF16, st.cuda_stream), "conv")
            b.record(st)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        ms = sorted(ms[1:])
        t = ms[len(ms) // 2]
        tot += t
        nb = 2 * B * hw * hw * (ci + co)
        print(json.dumps({"cin": ci, "cout": co, "hw": hw, "us": round(t * 1e3, 1), "GBps": round(nb / t / 1e6, 1),
                          "TFLOPs": round(2.0 * B * hw * hw * 9 * ci * co / t / 1e9, 1)}))
    print(json.dumps({"total_us": round(tot * 1e3, 1)}))


if __name__ == "__main__":
    main()
