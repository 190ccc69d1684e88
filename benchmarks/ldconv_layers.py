"""Per-kernel timing of the LDConv path at the 10 layer shapes of yolov8-LD-P2 (batch 64, 640x640):
CUDA events on the launching stream, inputs larger than L2 or L2 flushed between launches, algorithmic bytes of
SURVEY.md 8d.  Prints one JSON line per (layer, kernel, variant).
    python benchmarks/ldconv_layers.py [--batch 64] [--dtype bf16] [--iters 5] [--bwd]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import _lib  # noqa: E402

LAYERS = [  # layer, C, O, N, s, H (=W) at 640x640
    (0, 3, 16, 3, 2, 640), (1, 16, 32, 3, 2, 320), (3, 32, 64, 3, 2, 160), (5, 64, 128, 3, 2, 80), (8, 128, 64, 1, 1, 40),
    (10, 64, 64, 1, 1, 80), (13, 64, 32, 1, 1, 80), (15, 32, 32, 1, 1, 160), (18, 32, 32, 3, 2, 160), (21, 64, 64, 3, 2, 80)]


def timed(fn, iters, flush):
    st = torch.cuda.current_stream()
    ms = []
    for _ in range(iters + 1):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        fn()
        b.record(st)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms = sorted(ms[1:])
    return ms[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--sigma", type=float, default=0.5, help="std of the synthetic offsets in pixels")
    ap.add_argument("--scale", type=int, default=1, help="image scale: 2 = 1280x1280 (BASELINE config 5)")
    ap.add_argument("--bwd", action="store_true")
    args = ap.parse_args()
    L = _lib.load()
    dev = torch.device("cuda", 0)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    dt = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
    e = 2 if dtype == torch.bfloat16 else 4
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    st = torch.cuda.current_stream().cuda_stream
    B = args.batch
    for (li, C, O, N, s, H) in LAYERS:
        H = H * args.scale
        W = H
        h = w = (H - 1) // s + 1
        M, K = B * h * w, N * C
        g = torch.Generator(device=dev).manual_seed(li)
        x = torch.randn((B, H, W, C), device=dev, generator=g).to(dtype)
        w_off = torch.randn((3, 3, C, 2 * N), device=dev, generator=g) * 0.05
        b_off = torch.randn((2 * N,), device=dev, generator=g) * 0.1
        off = torch.randn((B, h, w, 2 * N), device=dev, generator=g) * args.sigma
        pn = torch.tensor(_lib.p_n_table(N), dtype=torch.int32, device=dev)
        operand = torch.empty((M, K), device=dev, dtype=dtype)
        wt = (torch.randn((O, K), device=dev, generator=g) * 0.1).to(dtype)
        scale = torch.ones(O, device=dev)
        shift = torch.zeros(O, device=dev)
        out = torch.empty((M, O), device=dev, dtype=dtype)
        off_out = torch.empty_like(off)

        def rec(kernel, variant, ms, nbytes, flops=0.0):
            print(json.dumps({"layer": li, "C": C, "O": O, "N": N, "s": s, "H": H, "kernel": kernel, "variant": variant,
                              "us": round(ms * 1e3, 1), "MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / ms / 1e6, 1),
                              "TFLOPs": round(flops / ms / 1e9, 2)}), flush=True)

        bytes_off = e * B * C * H * W + 4 * B * 2 * N * h * w
        ms = timed(lambda: _lib.check(L.ldconv_offset_conv_fwd(x.data_ptr(), w_off.data_ptr(), b_off.data_ptr(),
                                                               off_out.data_ptr(), B, C, H, W, N, s, dt, st)), args.iters, flush)
        rec("offset_conv_fwd", "ffma", ms, bytes_off, 2.0 * M * 9 * C * 2 * N)

        if L.ldconv_offset_conv_tc_supported(C, N, s, dt):
            w_tc = w_off.permute(3, 0, 1, 2).reshape(2 * N, 9 * C).to(dtype).contiguous()
            ms = timed(lambda: _lib.check(L.ldconv_offset_conv_tc_fwd(x.data_ptr(), w_tc.data_ptr(), b_off.data_ptr(),
                                                                      off_out.data_ptr(), B, C, H, W, N, s, dt, st)),
                       args.iters, flush)
            rec("offset_conv_fwd", "tcgen05", ms, bytes_off, 2.0 * M * 9 * C * 2 * N)

        if s == 2 and L.ldconv_offset_conv_s2d_supported(C, N, H, W, dt):
            from experiment_yolo_b200.ldconv import _prepare, base_grid
            pw = (w_off.permute(3, 2, 0, 1).contiguous()).to(dtype)          # (2N,C,3,3)
            pr = _prepare(pw, b_off, torch.zeros(8, C, N, 1, device=dev), base_grid(N), dtype, False)
            ms = timed(lambda: _lib.check(L.ldconv_offset_conv_s2d_fwd(x.data_ptr(), pr.w_off_s2d.data_ptr(), b_off.data_ptr(),
                                                                       off_out.data_ptr(), B, C, H, W, N, dt, st)),
                       args.iters, flush)
            rec("offset_conv_fwd", "tcgen05_s2d", ms, bytes_off, 2.0 * M * 9 * C * 2 * N)

        bytes_g = e * B * C * H * W + 4 * B * 2 * N * h * w + e * M * K
        for variant, direct in (("tma_tile", 0), ("direct", 1)):
            L.ldconv_set_flag(_lib.FLAG_GATHER_DIRECT, direct)
            ms = timed(lambda: _lib.check(L.ldconv_gather_fwd(x.data_ptr(), off.data_ptr(), pn.data_ptr(), operand.data_ptr(),
                                                              None, None, B, C, H, W, N, s, dt, st)), args.iters, flush)
            rec("gather_fwd", variant, ms, bytes_g)
        L.ldconv_set_flag(_lib.FLAG_GATHER_DIRECT, 0)

        bytes_mm = e * (M * K + O * K + M * O)
        for variant, ffma in (("auto", 0), ("ffma", 1)):
            L.ldconv_set_flag(_lib.FLAG_FORCE_FFMA, ffma)
            ms = timed(lambda: _lib.check(L.ldconv_gemm_fwd(operand.data_ptr(), wt.data_ptr(), scale.data_ptr(),
                                                            shift.data_ptr(), out.data_ptr(), None, None, None, M, K, O,
                                                            _lib.ACT_SILU, dt, st)), args.iters, flush)
            impl = L.ldconv_last_impl()
            rec("gemm_fwd", "tcgen05" if impl == _lib.IMPL_TCGEN05 else "ffma", ms, bytes_mm, 2.0 * M * K * O)
            if impl == _lib.IMPL_FFMA:
                break
        L.ldconv_set_flag(_lib.FLAG_FORCE_FFMA, 0)

        if L.ldconv_gather_gemm_supported(B, C, H, W, N, s, O, O, dt):
            bytes_gg = e * B * C * H * W + 4 * B * 2 * N * h * w + e * M * O
            ms = timed(lambda: _lib.check(L.ldconv_gather_gemm_fwd(x.data_ptr(), off.data_ptr(), pn.data_ptr(), wt.data_ptr(),
                                                                   scale.data_ptr(), shift.data_ptr(), out.data_ptr(), O, B, C, H,
                                                                   W, N, s, O, _lib.ACT_SILU, dt, st)), args.iters, flush)
            rec("gather_gemm_fwd", "tcgen05", ms, bytes_gg, 2.0 * M * K * O)

        if L.ldconv_fused_supported(B, C, H, W, N, s, O, dt):
            bytes_f = e * B * C * H * W + e * M * O
            ms = timed(lambda: _lib.check(L.ldconv_fused_fwd(x.data_ptr(), w_off.data_ptr(), b_off.data_ptr(), pn.data_ptr(),
                                                             wt.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(),
                                                             None, B, C, H, W, N, s, O, _lib.ACT_SILU, dt, st)), args.iters, flush)
            rec("fused_fwd", "tcgen05" if L.ldconv_last_impl() == _lib.IMPL_TCGEN05 else "smallc", ms, bytes_f,
                2.0 * M * K * O + 2.0 * M * 9 * C * 2 * N)

        if args.bwd:
            gop = torch.randn((M, K), device=dev, generator=g).to(dtype)
            grad_x = torch.zeros((B, H, W, C), device=dev)
            grad_off = torch.empty_like(off)
            bytes_s = e * M * K + e * B * C * H * W + 4 * B * 2 * N * h * w + 4 * B * C * H * W + 4 * B * 2 * N * h * w
            ms = timed(lambda: _lib.check(L.ldconv_gather_bwd(gop.data_ptr(), x.data_ptr(), off.data_ptr(), pn.data_ptr(),
                                                              grad_x.data_ptr(), grad_off.data_ptr(), B, C, H, W, N, s, dt, st)),
                       args.iters, flush)
            rec("gather_bwd", "atomics", ms, bytes_s)
            if dtype == torch.bfloat16 and L.ldconv_bwd_acc16_supported(B, C, H, W, N, s):
                grad_x16 = torch.zeros((B, H, W, C), device=dev, dtype=torch.bfloat16)
                bytes_16 = e * M * K + e * B * C * H * W + 4 * B * 2 * N * h * w + 2 * B * C * H * W + 4 * B * 2 * N * h * w
                ms = timed(lambda: _lib.check(L.ldconv_gather_bwd_acc16(gop.data_ptr(), x.data_ptr(), off.data_ptr(), pn.data_ptr(),
                                                                        grad_x16.data_ptr(), grad_off.data_ptr(), B, C, H, W, N, s, st)),
                           args.iters, flush)
                rec("gather_bwd", "atomics_bf16_accumulator", ms, bytes_16)
                del grad_x16
            gpre = torch.randn((M, O), device=dev, generator=g).to(dtype)
            gw = torch.zeros((O, K), device=dev)
            ms = timed(lambda: _lib.check(L.ldconv_gemm_bwd_weight(gpre.data_ptr(), operand.data_ptr(), gw.data_ptr(), M, K, O,
                                                                   dt, st)), args.iters, flush)
            rec("gemm_bwd_weight", "ffma", ms, e * (M * K + M * O), 2.0 * M * K * O)
        del x, operand, out, off, off_out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
