"""First LDConv layer (3 -> 16, num_param 3, stride 2, 640x640, batch 64, bf16) through ldconv_fused_fwd: CUDA events on the
launching stream, two alternating 157 MB inputs (larger than L2).  A/B through the library's environment switches:
    LDCONV_L0_ROWS=0|1 (thread-per-pixel kernel | rows kernel)  LDCONV_L0_MINB=5|6 (register budget of the rows kernel)
    python benchmarks/l0_ab.py [--batch 64] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--hw", type=int, default=640)
    args = ap.parse_args()
    L = _lib.load()
    dev = torch.device("cuda", 0)
    B, C, O, N, s, H = args.batch, 3, 16, 3, 2, args.hw
    W, h, w = H, H // 2, H // 2
    g = torch.Generator(device=dev).manual_seed(0)
    xs = [torch.randn((B, H, W, C), device=dev, generator=g).bfloat16() for _ in range(2)]
    w_off = torch.randn((3, 3, C, 2 * N), device=dev, generator=g) * 0.05
    b_off = torch.randn((2 * N,), device=dev, generator=g) * 0.1
    pn = torch.tensor(_lib.p_n_table(N), dtype=torch.int32, device=dev)
    wt = (torch.randn((O, N * C), device=dev, generator=g) * 0.1).bfloat16()
    scale, shift = torch.ones(O, device=dev), torch.zeros(O, device=dev)
    out = torch.empty((B, h, w, O), device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream()

    def run(x):
        _lib.check(L.ldconv_fused_fwd(x.data_ptr(), w_off.data_ptr(), b_off.data_ptr(), pn.data_ptr(), wt.data_ptr(),
                                      scale.data_ptr(), shift.data_ptr(), out.data_ptr(), None, B, C, H, W, N, s, O,
                                      _lib.ACT_SILU, _lib.BF16, st.cuda_stream), "ldconv_fused_fwd")

    for k in range(3):
        run(xs[k & 1])
    torch.cuda.synchronize()
    us = []
    for k in range(args.iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        run(xs[k & 1])
        b.record(st)
        torch.cuda.synchronize()
        us.append(a.elapsed_time(b) * 1e3)
    us.sort()
    med = us[len(us) // 2]
    nbytes = xs[0].numel() * 2 + out.numel() * 2
    print(json.dumps({"kernel": "ldconv_fused_fwd layer 0", "rows_kernel": os.environ.get("LDCONV_L0_ROWS", "1"),
                      "minb": os.environ.get("LDCONV_L0_MINB", "6"), "const_bank": os.environ.get("LDCONV_L0_CONST", "1"), "batch": B, "hw": H, "us_median": round(med, 1),
                      "us_min": round(us[0], 1), "MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / med / 1e3, 1),
                      "checksum": float(out.float().abs().mean())}))


if __name__ == "__main__":
    main()
