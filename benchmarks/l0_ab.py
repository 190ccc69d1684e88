"""First layer (3 -> 16, num_param 3, stride 2, 640 x 640, bf16) through ldconv_fused_fwd: tensor-core kernel (variant 0) against the
CUDA-core rows kernel (variant 1); L2 flushed between timed launches; also the max difference of the two outputs.
    python benchmarks/l0_ab.py [--batch 64] [--iters 7]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import experiment_yolo_b200 as E  # noqa: E402
from experiment_yolo_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--trace", action="store_true", help="clock64 stamps of CTA 0, tile iterations 2..5 of the tensor-core kernel")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    L = _lib.load()
    B, C, H, W, N, s, O = args.batch, 3, args.size, args.size, 3, 2, 16
    torch.manual_seed(0)
    mod = E.LDConv(C, O, N, s)
    with torch.no_grad():
        mod.p_conv.weight.normal_(0, 0.05)
    mod = mod.to(dev).bfloat16().eval()
    pr = mod._prepared(torch.bfloat16, False)
    x = torch.rand((B, H, W, C), device=dev).bfloat16()
    h, w = H // 2, W // 2
    scale = torch.rand(O, device=dev) + 0.5
    shift = torch.randn(O, device=dev) * 0.1
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    outs = {}
    for variant, name in ((0, "tcgen05"), (1, "rows_ffma")):
        L.ldconv_debug_l0_variant(variant)
        out = torch.empty((B, h, w, O), device=dev, dtype=torch.bfloat16)
        run = lambda: _lib.check(L.ldconv_fused_fwd(x.data_ptr(), pr.w_off.data_ptr(), pr.b_off.data_ptr(), pr.pn.data_ptr(), pr.wt.data_ptr(),
                                                    scale.data_ptr(), shift.data_ptr(), out.data_ptr(), None, B, C, H, W, N, s, O,
                                                    _lib.ACT_SILU, _lib.BF16, st), "ldconv_fused_fwd")
        for _ in range(2):
            run()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        nbytes = 2 * B * H * W * C + 2 * B * h * w * O
        outs[name] = out.float()
        print(json.dumps({"variant": name, "us_median": round(ts[len(ts) // 2], 1), "us_min": round(ts[0], 1), "MB": round(nbytes / 1e6, 1),
                          "GBps": round(nbytes / ts[len(ts) // 2] / 1e3, 1)}), flush=True)
    L.ldconv_debug_l0_variant(0)
    if args.trace:
        buf = torch.zeros(64, dtype=torch.int64, device=dev)
        L.ldconv_debug_l0_trace(buf.data_ptr())
        out = torch.empty((B, h, w, O), device=dev, dtype=torch.bfloat16)
        _lib.check(L.ldconv_fused_fwd(x.data_ptr(), pr.w_off.data_ptr(), pr.b_off.data_ptr(), pr.pn.data_ptr(), pr.wt.data_ptr(),
                                      scale.data_ptr(), shift.data_ptr(), out.data_ptr(), None, B, C, H, W, N, s, O,
                                      _lib.ACT_SILU, _lib.BF16, st), "ldconv_fused_fwd")
        torch.cuda.synchronize()
        L.ldconv_debug_l0_trace(None)
        print("grid", int(buf[63]))
        t = [int(v) for v in buf[:63].cpu().tolist() if v]
        names = ["tile start", "A1 stored", "barrier 1", "offsets ready (mbar)", "offsets in registers", "A2 stored", "barrier 2", "accumulator ready (mbar)"]
        for k in range(1, len(t)):
            print(f"{names[k % 8]:28s} +{t[k] - t[k - 1]:6d} cycles")
    d = (outs["tcgen05"] - outs["rows_ffma"]).abs()
    print(json.dumps({"max_abs_diff": float(d.max()), "rel_l2": float(d.norm() / outs["rows_ffma"].norm()),
                      "frac_elements_differing_by_more_than_0.02": float((d > 0.02).float().mean())}))


if __name__ == "__main__":
    main()
