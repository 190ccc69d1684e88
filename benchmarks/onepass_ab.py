"""A/B of the LDConv inference forward at the yolov8-LD-P2 layer shapes (batch 64, 640x640 -> SURVEY.md App. B):
  two-kernel path   ldconv_offset_conv_{s2d,tc}_fwd  +  ldconv_gather_gemm_fwd   (offsets round-trip through HBM, x read twice)
  one-pass kernel   ldconv_onepass_fwd                                          (x read once)
CUDA events on the launching stream, L2 flushed (256 MB memset) before every timed launch, median of --iters.
Algorithmic bytes of the one-pass kernel: e*B*C*H*W + e*M*O (x once + out).  One JSON line per layer.
    python benchmarks/onepass_ab.py [--batch 64] [--iters 7] [--scale 1] [--sigma 0.05]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import experiment_yolo_b200 as E  # noqa: E402
from experiment_yolo_b200 import _lib  # noqa: E402
from experiment_yolo_b200.ldconv import _folded_bn, offset_conv_nhwc  # noqa: E402

LAYERS = [(1, 16, 32, 3, 2, 320), (3, 32, 64, 3, 2, 160), (5, 64, 128, 3, 2, 80), (8, 128, 64, 1, 1, 40),
          (10, 64, 64, 1, 1, 80), (13, 64, 32, 1, 1, 80), (15, 32, 32, 1, 1, 160), (18, 32, 32, 3, 2, 160), (21, 64, 64, 3, 2, 80)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--sigma", type=float, default=0.05, help="std of p_conv.weight")
    ap.add_argument("--layers", default="")
    args = ap.parse_args()
    L = _lib.load()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    only = [int(v) for v in args.layers.split(",")] if args.layers else None
    B = args.batch

    def timed(fn):
        ms = []
        for _ in range(args.iters + 1):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            fn()
            b.record(st)
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        ms = sorted(ms[1:])
        return ms[len(ms) // 2]

    tot = {"two": 0.0, "one": 0.0}
    for (li, C, O, N, s, H) in LAYERS:
        if only and li not in only:
            continue
        H *= args.scale
        W = H
        h = w = (H - 1) // s + 1
        torch.manual_seed(li)
        mod = E.LDConv(C, O, N, s)
        with torch.no_grad():
            mod.p_conv.weight.normal_(0, args.sigma)
        mod = mod.to(dev).bfloat16().eval()
        x = torch.randn((B, H, W, C), device=dev).bfloat16()
        pr = mod._prepared(torch.bfloat16, False)
        scale, shift = _folded_bn(mod.conv[1], dev)
        out = torch.empty((B, h, w, O), device=dev, dtype=torch.bfloat16)
        out1 = torch.empty_like(out)
        off = offset_conv_nhwc(x, pr, N, s)
        sv = st.cuda_stream
        t_off = timed(lambda: offset_conv_nhwc(x, pr, N, s))
        t_gg = timed(lambda: _lib.check(L.ldconv_gather_gemm_fwd(x.data_ptr(), off.data_ptr(), pr.pn.data_ptr(), pr.wt.data_ptr(),
                                                                 scale.data_ptr(), shift.data_ptr(), out.data_ptr(), O, B, C, H, W, N, s, O,
                                                                 _lib.ACT_SILU, _lib.BF16, sv), "gg"))
        row = {"layer": li, "C": C, "O": O, "N": N, "s": s, "H": H, "offconv_us": round(t_off * 1e3, 1), "gg_us": round(t_gg * 1e3, 1),
               "two_kernel_us": round((t_off + t_gg) * 1e3, 1)}
        w_conv = pr.w_off_tc if s == 1 else pr.w_off_s2d
        if L.ldconv_onepass_supported(B, C, H, W, N, s, O, O, _lib.BF16):
            t_one = timed(lambda: _lib.check(L.ldconv_onepass_fwd(x.data_ptr(), w_conv.data_ptr(), pr.b_off.data_ptr(), pr.pn.data_ptr(),
                                                                  pr.wt.data_ptr(), scale.data_ptr(), shift.data_ptr(), out1.data_ptr(), O, None,
                                                                  B, C, H, W, N, s, O, _lib.ACT_SILU, _lib.BF16, sv), "one"))
            nb = 2 * B * C * H * W + 2 * B * h * w * O
            row.update({"onepass_us": round(t_one * 1e3, 1), "onepass_MB": round(nb / 1e6, 1), "onepass_GBps": round(nb / t_one / 1e6, 1),
                        "equal": bool(torch.equal(out, out1))})
            tot["two"] += t_off + t_gg
            tot["one"] += t_one
        print(json.dumps(row), flush=True)
    print(json.dumps({"total_two_kernel_us": round(tot["two"] * 1e3, 1), "total_onepass_us": round(tot["one"] * 1e3, 1)}))


if __name__ == "__main__":
    main()
