"""Timeline (clock64) of ldconv_onepass_kernel: worker thread 0 and the issuer thread of CTA 0, tile iterations 6..9.
    python benchmarks/trace_onepass.py --layer 15
tags (worker): 0 loop top, 1 offsets read from TMEM, 2 phase 1 done, 3 barrier A passed, 4 input tile / operand buffer acquired,
5 phase 2 done, 6 barrier B passed, 7 epilogue of the previous tile done; (issuer): 50 waiting for barrier B, 51 got it,
52 TMA issued, 53 main MMAs issued, 54 next input tile landed, 55 its offset-conv MMAs issued."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import experiment_yolo_b200 as E  # noqa: E402
from benchmarks.onepass_ab import LAYERS  # noqa: E402
from experiment_yolo_b200 import _lib  # noqa: E402
from experiment_yolo_b200.ldconv import _folded_bn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", type=int, default=15)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    L = _lib.load()
    dev = torch.device("cuda", 0)
    li, C, O, N, s, H = [l for l in LAYERS if l[0] == args.layer][0]
    W, B = H, args.batch
    h = w = (H - 1) // s + 1
    torch.manual_seed(li)
    mod = E.LDConv(C, O, N, s)
    with torch.no_grad():
        mod.p_conv.weight.normal_(0, 0.05)
    mod = mod.to(dev).bfloat16().eval()
    x = torch.randn((B, H, W, C), device=dev).bfloat16()
    pr = mod._prepared(torch.bfloat16, False)
    scale, shift = _folded_bn(mod.conv[1], dev)
    out = torch.empty((B, h, w, O), device=dev, dtype=torch.bfloat16)
    w_conv = pr.w_off_tc if s == 1 else pr.w_off_s2d
    buf = torch.zeros(2 * (2 * 64 + 2), device=dev, dtype=torch.int64)
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        if rep == 2:
            L.ldconv_debug_onepass_trace(buf.data_ptr())
        _lib.check(L.ldconv_onepass_fwd(x.data_ptr(), w_conv.data_ptr(), pr.b_off.data_ptr(), pr.pn.data_ptr(), pr.wt.data_ptr(),
                                        scale.data_ptr(), shift.data_ptr(), out.data_ptr(), O, None, B, C, H, W, N, s, O, _lib.ACT_SILU,
                                        _lib.BF16, st), "one")
    torch.cuda.synchronize()
    b = buf.cpu().tolist()
    ev = []
    for role in range(2):
        base = role * 130
        n = b[base + 128]
        ev += [(b[base + 2 * k + 1], b[base + 2 * k]) for k in range(n)]
    ev.sort()
    t0 = ev[0][0] if ev else 0
    print(f"layer {li}: C={C} O={O} N={N} s={s}")
    for t, tag in ev:
        print(f"{t - t0:8d}  it {tag // 100}  tag {tag % 100:2d}  {'issuer' if tag % 100 >= 50 else 'worker'}")


if __name__ == "__main__":
    main()
