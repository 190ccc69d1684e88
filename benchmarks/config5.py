"""BASELINE.json config 5: DEAL-YOLO-LD at 1280x1280 (high-resolution UAV tiles), batch 32, bf16, P2 head on a 320x320 map.

For every combination of SURVEY.md 8d's offset statistics -- p_conv.weight sigma in {0.05, 0.3} x p_conv.bias sigma in {0, 2, 8}
pixels -- one JSON line with
  * images/s of the full forward through engine.FusedDealYolo (CUDA-graph replay, device-resident batch, CUDA events);
  * per LDConv row: time of the one-pass kernel on that row's real input (L2 flushed) and its HALO-MISS FRACTION: the share
    of samples with a bilinear corner outside the TMA-staged input tile (16 x 8 output pixels + 2 input pixels of halo,
    csrc/ldconv_onepass_umma.cu), which the kernel serves from L2 instead of shared memory; and the share of samples that
    leave the IMAGE (the reference's clamp quirk).  Both are computed from the offsets the kernel itself used (off_out).
    python benchmarks/config5.py [--batch 32] [--img 1280] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import _lib, dealyolo, engine  # noqa: E402
from experiment_yolo_b200.ldconv import _folded_bn  # noqa: E402


def miss_fractions(off, pn, H, W, N, s):
    """off (B,h,w,2N) fp32 offsets -> (halo-miss fraction, out-of-image fraction) with the kernel's tile geometry"""
    B, h, w, _ = off.shape
    dev = off.device
    i = torch.arange(h, device=dev).view(1, h, 1, 1)
    j = torch.arange(w, device=dev).view(1, 1, w, 1)
    pr = (i * s + pn[:N].view(1, 1, 1, N)).float() + off[..., :N]
    pk = (j * s + pn[N:].view(1, 1, 1, N)).float() + off[..., N:]
    outside = (pr < 0) | (pr > H - 1) | (pk < 0) | (pk > W - 1)
    r0 = pr.floor().clamp(0, H - 1)
    r1 = (pr.floor() + 1).clamp(0, H - 1)
    k0 = pk.floor().clamp(0, W - 1)
    k1 = (pk.floor() + 1).clamp(0, W - 1)
    # tile origin / extent in input pixels (OPShape: stride 1: 21 x 13 from (i0 - 2, j0 - 2); stride 2: 36 x 20 from (2 i0 - 2, 2 j0 - 2))
    i0 = (i // 16) * 16
    j0 = (j // 8) * 8
    ro, ko = i0 * s - 2, j0 * s - 2
    rin, kin = (21, 13) if s == 1 else (36, 20)
    inside = (r0 >= ro) & (r1 < ro + rin) & (k0 >= ko) & (k1 < ko + kin)
    return float((~inside).float().mean()), float(outside.float().mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--img", type=int, default=1280)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    L = _lib.load()
    dev = torch.device("cuda", 0)
    B, IMG = args.batch, args.img
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream()
    xs = [torch.rand((B, 3, IMG, IMG), device=dev).bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(2)]
    for sigma in (0.05, 0.3):
        for bias_sigma in (0.0, 2.0, 8.0):
            model = dealyolo.DealYolo(nc=6)
            model.load_state_dict(dealyolo.seeded_state(model, 0, p_conv_sigma=sigma))
            g = torch.Generator().manual_seed(7)
            with torch.no_grad():
                for m in model.ldconv_layers():
                    m.p_conv.bias.copy_(torch.randn(m.p_conv.bias.shape, generator=g) * bias_sigma)
            model = dealyolo.channels_last_(model.to(dev).bfloat16().eval())
            run = engine.FusedDealYolo(model)
            static_x = xs[0].clone()
            with torch.inference_mode():
                side = torch.cuda.Stream()
                side.wait_stream(st)
                with torch.cuda.stream(side):
                    for _ in range(2):
                        run(static_x)
                st.wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    y, _ = run(static_x)
                for k in range(3):
                    static_x.copy_(xs[k & 1]); graph.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(args.steps):
                    static_x.copy_(xs[k & 1]); graph.replay()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                finite = bool(torch.isfinite(y.float()).all())
                # per LDConv row: the row's real input through the one-pass kernel with off_out
                feats = {}
                hooks = [m.register_forward_pre_hook(lambda mod, inp, i=m.i: feats.__setitem__(i, inp[0])) for m in model.ldconv_layers()]
                model(xs[0][:4])
                for hk in hooks:
                    hk.remove()
            rows = []
            for m in model.ldconv_layers():
                xin = feats[m.i]
                b4, C, H, W = xin.shape
                N, s, O = m.num_param, int(m.stride), m.conv[0].out_channels
                h, w = (H - 1) // s + 1, (W - 1) // s + 1
                pr = m._prepared(torch.bfloat16, False)
                w_conv = pr.w_off_tc if s == 1 else pr.w_off_s2d
                if w_conv is None or not L.ldconv_onepass_supported(b4, C, H, W, N, s, O, O, _lib.BF16):
                    continue
                scale, shift = _folded_bn(m.conv[1], dev)
                xh = xin.permute(0, 2, 3, 1).contiguous()
                off = torch.empty((b4, h, w, 2 * N), device=dev, dtype=torch.float32)
                out = torch.empty((b4, h, w, O), device=dev, dtype=torch.bfloat16)
                _lib.check(L.ldconv_onepass_fwd(xh.data_ptr(), w_conv.data_ptr(), pr.b_off.data_ptr(), pr.pn.data_ptr(), pr.wt.data_ptr(),
                                                scale.data_ptr(), shift.data_ptr(), out.data_ptr(), O, off.data_ptr(), b4, C, H, W, N, s, O,
                                                _lib.ACT_SILU, _lib.BF16, st.cuda_stream), "one")
                torch.cuda.synchronize()
                miss, oob = miss_fractions(off, pr.pn, H, W, N, s)
                rows.append({"layer": m.i, "C": C, "N": N, "s": s, "map": [H, W], "offset_abs_mean_px": round(float(off.abs().mean()), 2),
                             "halo_miss_frac": round(miss, 4), "out_of_image_frac": round(oob, 4)})
            print(json.dumps({"config": "DEAL-YOLO-LD full forward, %dx%d, batch %d, bf16" % (IMG, IMG, B), "p_conv_weight_sigma": sigma,
                              "p_conv_bias_sigma_px": bias_sigma, "ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1),
                              "finite": finite, "ldconv_rows": rows}), flush=True)
            del run, graph, model, y
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
