"""Run ONE C-ABI kernel of the LDConv path a few times at a yolov8-LD-P2 layer shape (for `ncu --set full -k regex:...`).
    python benchmarks/one_kernel.py --kernel gemm|gather|offset|fused --layer 1 [--batch 64] [--reps 3] [--direct]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from benchmarks.ldconv_layers import LAYERS  # noqa: E402
from experiment_yolo_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="gemm")
    ap.add_argument("--layer", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--direct", action="store_true")
    ap.add_argument("--ffma", action="store_true")
    ap.add_argument("--sigma", type=float, default=0.5)
    ap.add_argument("--cin", type=int, default=32, help="--kernel conv3x3 / conv1x1: input channels")
    ap.add_argument("--cout", type=int, default=32)
    ap.add_argument("--hw", type=int, default=160)
    ap.add_argument("--stride", type=int, default=1)
    args = ap.parse_args()
    L = _lib.load()
    dev = torch.device("cuda", 0)
    li, C, O, N, s, H = [l for l in LAYERS if l[0] == args.layer][0]
    W, B = H, args.batch
    h = w = (H - 1) // s + 1
    M, K = B * h * w, N * C
    dt, dtype = _lib.BF16, torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(li)
    x = torch.randn((B, H, W, C), device=dev, generator=g).to(dtype)
    w_off = torch.randn((3, 3, C, 2 * N), device=dev, generator=g) * 0.05
    b_off = torch.randn((2 * N,), device=dev, generator=g) * 0.1
    off = torch.randn((B, h, w, 2 * N), device=dev, generator=g) * args.sigma
    pn = torch.tensor(_lib.p_n_table(N), dtype=torch.int32, device=dev)
    operand = torch.randn((M, K), device=dev, generator=g).to(dtype)
    wt = (torch.randn((O, K), device=dev, generator=g) * 0.1).to(dtype)
    scale, shift = torch.ones(O, device=dev), torch.zeros(O, device=dev)
    out = torch.empty((M, O), device=dev, dtype=dtype)
    st = torch.cuda.current_stream().cuda_stream
    L.ldconv_set_flag(_lib.FLAG_GATHER_DIRECT, int(args.direct))
    L.ldconv_set_flag(_lib.FLAG_FORCE_FFMA, int(args.ffma))
    if args.kernel in ("conv3x3", "conv1x1"):
        ci, co, hw, cs = args.cin, args.cout, args.hw, args.stride
        ho = (hw - 1) // cs + 1
        xc = torch.randn((B, hw, hw, ci), device=dev, generator=g).to(dtype)
        kk = 9 if args.kernel == "conv3x3" else 1
        wc = (torch.randn((co, kk * ci), device=dev, generator=g) * 0.05).to(dtype)
        sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
        oc = torch.empty((B, ho, ho, co), device=dev, dtype=dtype)
    for _ in range(args.reps):
        if args.kernel == "conv3x3":
            _lib.check(L.ldconv_conv3x3_bn_act_fwd(xc.data_ptr(), ci, wc.data_ptr(), sc.data_ptr(), sh.data_ptr(), None, 0,
                                                   oc.data_ptr(), co, B, ci, hw, hw, co, cs, _lib.ACT_SILU, dt, st))
        elif args.kernel == "conv1x1":
            _lib.check(L.ldconv_conv1x1_bn_act_fwd(xc.data_ptr(), ci, wc.data_ptr(), sc.data_ptr(), sh.data_ptr(), None, 0,
                                                   oc.data_ptr(), co, B * hw * hw, ci, co, _lib.ACT_SILU, dt, st))
        elif args.kernel == "gemm":
            _lib.check(L.ldconv_gemm_fwd(operand.data_ptr(), wt.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(),
                                         None, None, None, M, K, O, _lib.ACT_SILU, dt, st))
        elif args.kernel == "gather":
            _lib.check(L.ldconv_gather_fwd(x.data_ptr(), off.data_ptr(), pn.data_ptr(), operand.data_ptr(), None, None, B, C,
                                           H, W, N, s, dt, st))
        elif args.kernel == "offset":
            _lib.check(L.ldconv_offset_conv_fwd(x.data_ptr(), w_off.data_ptr(), b_off.data_ptr(), off.data_ptr(), B, C, H, W,
                                                N, s, dt, st))
        elif args.kernel == "offset_tc":
            w_tc = w_off.permute(3, 0, 1, 2).reshape(2 * N, 9 * C).to(dtype).contiguous()
            _lib.check(L.ldconv_offset_conv_tc_fwd(x.data_ptr(), w_tc.data_ptr(), b_off.data_ptr(), off.data_ptr(), B, C, H, W,
                                                   N, s, dt, st))
        elif args.kernel == "gg":
            _lib.check(L.ldconv_gather_gemm_fwd(x.data_ptr(), off.data_ptr(), pn.data_ptr(), wt.data_ptr(), scale.data_ptr(),
                                                shift.data_ptr(), out.data_ptr(), O, B, C, H, W, N, s, O, _lib.ACT_SILU, dt, st))
        elif args.kernel == "fused":
            _lib.check(L.ldconv_fused_fwd(x.data_ptr(), w_off.data_ptr(), b_off.data_ptr(), pn.data_ptr(), wt.data_ptr(),
                                          scale.data_ptr(), shift.data_ptr(), out.data_ptr(), None, B, C, H, W, N, s, O,
                                          _lib.ACT_SILU, dt, st))
        torch.cuda.synchronize()
    print("ok", args.kernel, args.layer)


if __name__ == "__main__":
    main()
