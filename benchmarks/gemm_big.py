"""The (N,1)-conv GEMM of LDConv at the large-K end of BASELINE.json config 2 (C = 256, num_param 9 -> K = 2304, O = 256; M = 64 x 40 x 40
rows), bf16, through ldconv_gemm_fwd: TFLOP/s from CUDA events; run under `ncu --set full -k regex:umma_gemm` for the tensor-pipe
counters (profiles/r2_ncu_gemm_k2304.txt).
    python benchmarks/gemm_big.py [--iters 10]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    L = _lib.load()
    M, K, O = 64 * 40 * 40, 2304, 256
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn((M, K), device=dev, generator=g).bfloat16()
    wt = (torch.randn((O, K), device=dev, generator=g) * 0.02).bfloat16()
    scale, shift = torch.ones(O, device=dev), torch.zeros(O, device=dev)
    out = torch.empty((M, O), device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: _lib.check(L.ldconv_gemm_fwd(a.data_ptr(), wt.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(), None, None,
                                               None, M, K, O, _lib.ACT_SILU, _lib.BF16, st), "ldconv_gemm_fwd")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print(json.dumps({"kernel": "ldconv_gemm_fwd (umma_gemm_kernel)", "M": M, "K": K, "O": O, "us": round(ms * 1e3, 1),
                      "TFLOPs": round(2.0 * M * K * O / ms / 1e9, 1), "GBps": round(2.0 * (M * K + O * K + M * O) / ms / 1e6, 1)}))


if __name__ == "__main__":
    main()
