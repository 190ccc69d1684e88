"""Narrow 1x1 conv (rows = 64 x 160 x 160 pixels) through the plain tcgen05 GEMM and the pixel-packed one (P = 2, 4): CUDA events,
two alternating inputs, median.    python benchmarks/gemm_pack_ab.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import _lib  # noqa: E402


def main():
    L = _lib.load()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream()
    for (K, O, rows) in [(32, 32, 64 * 160 * 160), (48, 32, 64 * 160 * 160), (64, 64, 64 * 160 * 160), (64, 64, 64 * 80 * 80)]:
        g = torch.Generator(device=dev).manual_seed(K)
        xs = [torch.randn((rows, K), device=dev, generator=g).bfloat16() for _ in range(2)]
        wt = (torch.randn((O, K), device=dev, generator=g) * 0.1).bfloat16()
        sc, sh = torch.ones(O, device=dev), torch.zeros(O, device=dev)
        out = torch.empty((rows, O), device=dev, dtype=torch.bfloat16)
        for P in (1, 2, 4):
            if P * O > 256:
                continue
            wp = torch.block_diag(*([wt.float()] * P)).bfloat16().contiguous()
            scp, shp = sc.repeat(P).contiguous(), sh.repeat(P).contiguous()

            def run(x):
                if P == 1:
                    _lib.check(L.ldconv_conv1x1_bn_act_fwd(x.data_ptr(), K, wt.data_ptr(), sc.data_ptr(), sh.data_ptr(), None, 0,
                                                           out.data_ptr(), O, rows, K, O, _lib.ACT_SILU, _lib.BF16, st.cuda_stream))
                else:
                    _lib.check(L.ldconv_conv1x1_bn_act_packed_fwd(x.data_ptr(), wp.data_ptr(), scp.data_ptr(), shp.data_ptr(),
                                                                  out.data_ptr(), O, rows, K, O, P, _lib.ACT_SILU, _lib.BF16,
                                                                  st.cuda_stream))
            for k in range(3):
                run(xs[k & 1])
            torch.cuda.synchronize()
            us = []
            for k in range(15):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                run(xs[k & 1])
                b.record(st)
                torch.cuda.synchronize()
                us.append(a.elapsed_time(b) * 1e3)
            us.sort()
            nbytes = rows * (K + O) * 2
            print(json.dumps({"K": K, "O": O, "rows": rows, "P": P, "us": round(us[len(us) // 2], 1),
                              "GBps": round(nbytes / us[len(us) // 2] / 1e3, 1)}))


if __name__ == "__main__":
    main()
