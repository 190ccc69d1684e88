"""BASELINE.json config 2: single LDConv layer sweep, C = O in {16..256}, num_param in {5, 9}, P2-P4 maps (160/80/40) at
stride 1 and 2, batch 64, bf16 and fp32, forward (inference path) and forward+backward (training path) through the module.
Also times the (N,1)-conv GEMM alone (tcgen05) and reports TFLOP/s -- the tensor-pipe figure of the north star is only
meaningful at the large-K end (C=256, N=9 -> K=2304, O=256).
    python benchmarks/sweep.py [--batch 64] [--quick] > profiles/sweep.jsonl
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import experiment_yolo_b200 as E  # noqa: E402
from experiment_yolo_b200 import _lib  # noqa: E402


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sorted(ms)[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    L = _lib.load()
    Cs = [16, 64, 256] if args.quick else [16, 32, 64, 128, 256]
    Ns = [9] if args.quick else [5, 9]
    maps = [(160, 1), (40, 2)] if args.quick else [(160, 1), (160, 2), (80, 1), (80, 2), (40, 1), (40, 2)]
    for dtype in (torch.bfloat16, torch.float32):
        for C in Cs:
            for N in Ns:
                for (H, s) in maps:
                    B = args.batch
                    torch.manual_seed(0)
                    mod = E.LDConv(C, C, N, s).to(dev)
                    with torch.no_grad():
                        mod.p_conv.weight.normal_(0, 0.05)
                    mod = mod.to(dtype)
                    x = torch.randn(B, C, H, H, device=dev).to(dtype).contiguous(memory_format=torch.channels_last)
                    h = (H - 1) // s + 1
                    M, K, e = B * h * h, N * C, x.element_size()
                    rec = {"dtype": str(dtype).replace("torch.", ""), "C": C, "O": C, "N": N, "H": H, "s": s, "B": B, "M": M, "K": K}
                    mod.eval()
                    with torch.no_grad():
                        rec["fwd_ms"] = round(timed(lambda: mod(x)), 4)
                    mod.train()
                    xg = x.clone().requires_grad_(True)
                    gout = torch.randn(B, C, h, h, device=dev).to(dtype).contiguous(memory_format=torch.channels_last)

                    def step():
                        mod.zero_grad(set_to_none=True)
                        xg.grad = None
                        mod(xg).backward(gout)
                    rec["fwd_bwd_ms"] = round(timed(step), 4)
                    # the GEMM alone
                    a = torch.randn(M, K, device=dev).to(dtype)
                    wt = (torch.randn(C, K, device=dev) * 0.05).to(dtype)
                    out = torch.empty(M, C, device=dev, dtype=dtype)
                    sc, sh = torch.ones(C, device=dev), torch.zeros(C, device=dev)
                    dt = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
                    st = torch.cuda.current_stream().cuda_stream
                    ms = timed(lambda: _lib.check(L.ldconv_gemm_fwd(a.data_ptr(), wt.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                                                                    out.data_ptr(), None, None, None, M, K, C, 1, dt, st)))
                    rec["gemm_ms"] = round(ms, 4)
                    rec["gemm_impl"] = "tcgen05" if L.ldconv_last_impl() == _lib.IMPL_TCGEN05 else "ffma"
                    rec["gemm_tflops"] = round(2.0 * M * K * C / ms / 1e9, 2)
                    rec["gemm_GBps"] = round(e * (M * K + C * K + M * C) / ms / 1e6, 1)
                    rec["gemm_flop_per_byte"] = round(2.0 * K * C / (e * (K + C)), 1)
                    print(json.dumps(rec), flush=True)
                    del mod, x, xg, gout, a, wt, out
                    torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
