"""rel-L2 error of the bf16-accumulator scatter (ldconv_gather_bwd_acc16) against the fp32-accumulator kernel on the same inputs,
over offset scales (0.5 px = the benchmark regime ... 60 px = most samples clamp onto the borders and pile up there).
    python benchmarks/acc16_error.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import _lib  # noqa: E402


def main():
    L = _lib.load()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=dev).manual_seed(0)
    for (C, N, s, H, W, B) in [(16, 3, 2, 320, 320, 4), (32, 3, 2, 160, 160, 4), (64, 1, 1, 80, 80, 4), (64, 3, 2, 80, 80, 4), (16, 3, 2, 37, 53, 2)]:
        h, w = (H - 1) // s + 1, (W - 1) // s + 1
        M = B * h * w
        x = torch.randn((B, H, W, C), device=dev, generator=g).bfloat16()
        gop = torch.randn((M, N * C), device=dev, generator=g).bfloat16()
        pn = torch.tensor(_lib.p_n_table(N), dtype=torch.int32, device=dev)
        for scale in (0.5, 3.0, 10.0, 60.0):
            off = torch.randn((B, h, w, 2 * N), device=dev, generator=g) * scale
            gx32 = torch.zeros((B, H, W, C), device=dev)
            go32 = torch.empty_like(off)
            _lib.check(L.ldconv_gather_bwd(gop.data_ptr(), x.data_ptr(), off.data_ptr(), pn.data_ptr(), gx32.data_ptr(), go32.data_ptr(),
                                           B, C, H, W, N, s, _lib.BF16, st), "ldconv_gather_bwd")
            gx16 = torch.zeros((B, H, W, C), device=dev, dtype=torch.bfloat16)
            go16 = torch.full_like(off, 7.0)
            _lib.check(L.ldconv_gather_bwd_acc16(gop.data_ptr(), x.data_ptr(), off.data_ptr(), pn.data_ptr(), gx16.data_ptr(),
                                                 go16.data_ptr(), B, C, H, W, N, s, st), "ldconv_gather_bwd_acc16")
            torch.cuda.synchronize()
            rel = float((gx16.float() - gx32).norm() / gx32.norm())
            one = float((gx32.bfloat16().float() - gx32).norm() / gx32.norm())
            reloff = float((go16 - go32).abs().max() / go32.abs().max().clamp_min(1.0))
            print(json.dumps({"C": C, "N": N, "s": s, "H": H, "offset_sigma_px": scale, "grad_x_rel_l2_bf16_accumulator": round(rel, 5),
                              "grad_x_rel_l2_single_bf16_rounding": round(one, 5), "grad_off_max_rel": reloff}), flush=True)


if __name__ == "__main__":
    main()
