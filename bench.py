#!/usr/bin/env python
"""bench.py -- DEAL-YOLO-LD 640x640 images/s on B200 (BASELINE.json metric), with the LDConv roofline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one full forward of DEAL-YOLO-LD (cfg/deal-yolo-ld-p2.yaml = the reference's yolov8-LD-P2.yaml, 918,304
parameters, random-init seeded weights) over one synthetic batch of 64 bf16 640x640 images per GPU, inference mode,
channels_last; the batch is sharded over ranks with no collective (every LDConv sample depends on its own image only).
  value     images/s with the batch already resident in HBM (CUDA-graph replay of the forward), max over ranks
  e2e       images/s through the public call `engine.PipelinedPredictor.submit()/result()` with HOST buffers: every step copies its
            uint8 batch from pinned host memory, normalises on the device, runs the forward and reads the result back to the host
  roofline  the LDConv gather+GEMM kernel at its largest launch of the step (`roofline_scatter`: the backward scatter of the training
            path with the bf16 accumulator at the same shapes; and `roofline_gather`: the stand-alone gather
            kernel): algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the eager CPU port of the reference path (oracle/ldconv_torch_port.py inside the same
            graph), all host threads, on a bounded sample (batch 8) of the same workload
Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

IMG = 640
PER_GPU_BATCH = 64
CPU_SAMPLE_BATCH = 8
NC = 6


# ---------------------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a thread every
    ~2 ms (a 64-image step is ~4 ms, `nvidia-smi -lms` cannot sample that fast); nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.stop, self.nvml = index, [], None, False, None
        self.sm, self.mx, self.reasons = [], 0.0, set()

    def _poll_nvml(self):
        n, h = self.nvml, self.handle
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop = True
        if self.nvml is not None:
            self.t.join(timeout=1)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = list(self.sm), self.mx, set(self.reasons)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------------ CPU baseline
def cpu_port_images_per_s(batch: int, iters: int, warmup: int = 1):
    """Eager CPU port of the reference path inside the same graph, fp32, all host threads."""
    import torch
    from experiment_yolo_b200 import dealyolo
    from oracle.ldconv_torch_port import LDConvTorchPort
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = dealyolo.DealYolo(nc=NC, ldconv_cls=LDConvTorchPort)
    model.load_state_dict(dealyolo.seeded_state(model, 0))
    model.eval()
    x = torch.rand(batch, 3, IMG, IMG, generator=torch.Generator().manual_seed(0))
    times = []
    with torch.inference_mode():
        for _ in range(warmup):
            model(x)
        for _ in range(iters):
            t0 = time.perf_counter()
            model(x)
            times.append(time.perf_counter() - t0)
    return batch / min(times), batch / (sum(times) / len(times)), sum(times) / len(times), cores


def gpu_eager_leg(dtype_name: str, batch: int, iters: int, local: int):
    """one dtype of the GPU eager baseline, in its OWN process (see gpu_eager_images_per_s); prints one JSON object"""
    import torch
    from experiment_yolo_b200 import dealyolo
    from oracle.ldconv_torch_port import LDConvTorchPort
    dtype = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[dtype_name]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    model = dealyolo.DealYolo(nc=NC, ldconv_cls=LDConvTorchPort)
    model.load_state_dict(dealyolo.seeded_state(model, 0))
    model = dealyolo.channels_last_(model.to(dev).to(dtype).eval())
    x = torch.rand((batch, 3, IMG, IMG), device=dev).to(dtype).contiguous(memory_format=torch.channels_last)
    with torch.inference_mode():
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            model(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    _emit({"value": round(batch / (ms / 1e3), 1), "ms_per_step": round(ms, 2)})
    return 0


def gpu_eager_images_per_s(local: int, batch: int, iters: int = 3):
    """SURVEY.md 8d's second, fairer baseline: the eager reference algorithm ON THE GPU -- the same graph with the eager
    port of the reference LDConv (its ATen op sequence, host-built p_0 and int64 gathers included), cuDNN convs around it,
    get_FPS.py:63-88 protocol (warm-up, timed loop, synchronize) under inference_mode, in fp32 and in the reference's own
    low-precision style (coordinates computed in the low precision too; timing only, its numbers are not used).  Each dtype
    runs in a subprocess: the low-precision coordinate arithmetic of the reference can produce an out-of-range gather index
    (a device-side assert, which would poison this process's CUDA context)."""
    out = {}
    for name in ("fp32", "bf16", "fp16"):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--eager-leg", name, "--eager-batch", str(batch),
                                "--eager-iters", str(iters), "--eager-device", str(local)], capture_output=True, text=True, timeout=240)
            if r.returncode == 0 and r.stdout.strip():
                out[name] = json.loads(r.stdout.strip().splitlines()[-1])
            else:
                err = [l for l in r.stderr.splitlines() if "Error" in l or "assert" in l.lower()]
                out[name] = {"error": (err[-1] if err else f"exit code {r.returncode}")[:160]}
        except Exception as e:
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
    out.update({"unit": "images/s", "batch": batch, "iters": iters,
                "what": "eager port of the reference LDConv + torch/cuDNN graph on this GPU, inference_mode, channels_last"})
    return out


def lib_sha16() -> str:
    """Identity of the library build: sha256 over its SOURCES (csrc/*.cu, *.cuh, Makefile, include/*.h, in name order).  The binary
    itself is not reproducible bit for bit (a clean rebuild of the same sources hashes differently), so the stamp that ties an ncu
    capture to the library being timed is taken over what the binary is built from."""
    import glob
    import hashlib
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(ROOT, "experiment_yolo_b200", "csrc", "*.cu")) +
                   glob.glob(os.path.join(ROOT, "experiment_yolo_b200", "csrc", "*.cuh")) +
                   glob.glob(os.path.join(ROOT, "experiment_yolo_b200", "csrc", "Makefile")) +
                   glob.glob(os.path.join(ROOT, "include", "*.h")))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def committed_traffic():
    """ncu DRAM byte counts captured on a build of the library (profiles/r2_ncu_traffic.json, written by
    scripts/ncu_traffic.py from `ncu --set full` reports); only used when its lib_sha16 equals the library being timed."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
    except Exception:
        return {}, "no profiles/r2_ncu_traffic.json"
    sha = lib_sha16()
    if t.get("lib_sha16") != sha:
        return {}, f"profiles/r2_ncu_traffic.json was captured on library {t.get('lib_sha16')}, this run times {sha}: not reported"
    return t, f"ncu --set full capture of this exact library build ({sha}), profiles/r2_ncu_traffic.json"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cpu_batch = args.cpu_batch or CPU_SAMPLE_BATCH       # --cpu-batch 1: BASELINE.json config 1 (benchmarks/config1.py)
    best, mean, sec, cores = cpu_port_images_per_s(cpu_batch, max(1, args.steps), max(1, min(args.warmup, 2)))
    sample = f"batch {cpu_batch} of 640x640 fp32 images per step, eager CPU port of the reference graph"
    line = {"impl": "reference", "metric": "images_per_sec", "value": round(mean, 3), "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.gpus), timed_batch=cpu_batch, engine="eager CPU port",
                           note=f"each step times a bounded sample of the workload: batch {cpu_batch} of the "
                                f"{PER_GPU_BATCH}-image batch (the CPU port's images/s at batch 64 and batch 8 agree, DESIGN.md 6)"),
            "cpu_baseline": {"value": round(mean, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(mean, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)
    return 0


def per_gpu_batch(n_gpus: int, scaling: str = "weak") -> int:
    """weak: 64 images per GPU (global 64 N); strong: SURVEY.md 8d config 3, one 64-image batch split 64/32/16/8"""
    return PER_GPU_BATCH if scaling == "weak" else max(1, PER_GPU_BATCH // n_gpus)


def n_input_buffers(batch: int) -> int:
    return max(2, -(-300_000_000 // (batch * 3 * IMG * IMG * 2)))


def workload_config(n_gpus: int, engine: str = "fused", scaling: str = "weak"):
    pgb = per_gpu_batch(n_gpus, scaling)
    return {"engine": engine, "workload": "DEAL-YOLO-LD (yolov8-LD-P2 graph, 10 LDConv + SSFF, nc=6) full forward, 640x640, inference",
            "global_batch": pgb * n_gpus, "per_gpu_batch": pgb, "imgsz": IMG,
            "layout": "channels_last", "parallelism": f"batch-sharded x{n_gpus}, no collective",
            "l2_policy": f"inputs_exceed_l2 ({pgb * 3 * IMG * IMG * 2 / 1e6:.0f} MB bf16 batch, {n_input_buffers(pgb)} rotating input buffers "
                         f"= {n_input_buffers(pgb) * pgb * 3 * IMG * IMG * 2 / 1e6:.0f} MB > 126 MB L2; the activations of one step are "
                         f"{6.5 * pgb / 64:.1f} GB of DRAM traffic)",
            "weights": "seeded random init (dealyolo.seeded_state(0)), p_conv.weight ~ N(0,0.05)"}


# --------------------------------------------------------------------------------------------------------------- GPU arm
def ldconv_roofline(model, x, peaks, iters: int):
    """Time the LDConv kernels of one step at this step's shapes with CUDA events on the launching stream (L2 flushed before
    every timed launch), feeding them the real layer inputs.  Kernels (SURVEY.md 8d formulas for the algorithmic bytes):
      one-pass     ldconv_onepass_fwd, THE kernel the inference step runs per LDConv row with C >= 16: x + out
                   (offsets and the resampled operand never reach HBM)
      two-kernel   the round-1 path for comparison: tensor-core offset conv (x + offsets) + ldconv_gather_gemm_fwd (x + offsets + out)
      gather       ldconv_gather_fwd, the stand-alone resampling kernel of the training path: x + offsets + operand
    `roofline` (the contract's key) is the one-pass kernel at its largest launch (layer 1: 16->32 channels, 3 samples,
    stride 2, 320x320 -> 160x160, batch 64); `traffic` is that launch's dram__bytes_read + dram__bytes_write from an
    `ncu --set full` capture of the same library build (profiles/r2_ncu_traffic.json), else null."""
    import torch
    from experiment_yolo_b200 import _lib
    from experiment_yolo_b200.ldconv import _folded_bn, offset_conv_nhwc
    L = _lib.load()
    feats = {}
    hooks = [m.register_forward_pre_hook(lambda mod, inp, i=m.i: feats.__setitem__(i, inp[0])) for m in model.ldconv_layers()]
    with torch.inference_mode():
        model(x)
    for h in hooks:
        h.remove()
    st = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, device=x.device, dtype=torch.uint8)      # > the 126 MB L2

    def timed(fn):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters + 1)]
        for a, b in ev:
            flush.zero_()            # L2 flush between the timed launches: every launch reads its input from HBM
            a.record(st)
            fn()
            b.record(st)
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev[1:])
        return ms[len(ms) // 2]

    per_layer, tot = [], {"one": [0.0, 0.0], "two": [0.0, 0.0], "gather": [0.0, 0.0], "scatter": [0.0, 0.0]}
    for m in model.ldconv_layers():
        xin = feats[m.i]
        B, C, H, W = xin.shape
        N, s, O = m.num_param, int(m.stride), m.conv[0].out_channels
        h, w = (H - 1) // s + 1, (W - 1) // s + 1
        e = xin.element_size()
        xh = xin.permute(0, 2, 3, 1).contiguous()
        pr = m._prepared(xin.dtype, False)
        scale, shift = _folded_bn(m.conv[1], xin.device)
        dt = _lib.BF16 if xin.dtype == torch.bfloat16 else _lib.F32
        sv = st.cuda_stream
        off = offset_conv_nhwc(xh, pr, N, s)
        row = {"layer": m.i, "C": C, "O": O, "N": N, "s": s, "hw": [h, w]}
        x_bytes, off_bytes, out_bytes = e * B * C * H * W, 4 * B * 2 * N * h * w, e * B * h * w * O
        operand = torch.empty((B * h * w, N * C), device=xin.device, dtype=xin.dtype)
        ms = timed(lambda: _lib.check(L.ldconv_gather_fwd(xh.data_ptr(), off.data_ptr(), pr.pn.data_ptr(), operand.data_ptr(),
                                                          None, None, B, C, H, W, N, s, dt, sv), "ldconv_gather_fwd"))
        nbytes = x_bytes + off_bytes + e * B * h * w * N * C
        row.update({"gather_MB": round(nbytes / 1e6, 1), "gather_us": round(ms * 1e3, 1), "gather_GBps": round(nbytes / ms / 1e6, 1)})
        if C >= 8:          # layer 0 (C = 3) runs the one-kernel small-C path in the step, not these kernels
            tot["gather"][0] += nbytes
            tot["gather"][1] += ms
        if dt == _lib.BF16 and L.ldconv_bwd_acc16_supported(B, C, H, W, N, s):
            # backward scatter of the training path (autograd of the four gathers): grad_operand + x + offsets in, grad_x (bf16
            # accumulator) + grad_offset out
            gop = torch.randn((B * h * w, N * C), device=xin.device, dtype=torch.float32).to(xin.dtype)
            gx = torch.zeros((B, H, W, C), device=xin.device, dtype=xin.dtype)
            goff = torch.empty_like(off)
            ms = timed(lambda: _lib.check(L.ldconv_gather_bwd_acc16(gop.data_ptr(), xh.data_ptr(), off.data_ptr(), pr.pn.data_ptr(),
                                                                    gx.data_ptr(), goff.data_ptr(), B, C, H, W, N, s, sv),
                                          "ldconv_gather_bwd_acc16"))
            nb = e * B * h * w * N * C + x_bytes + off_bytes + e * B * C * H * W + off_bytes
            row.update({"scatter_MB": round(nb / 1e6, 1), "scatter_us": round(ms * 1e3, 1), "scatter_GBps": round(nb / ms / 1e6, 1)})
            tot["scatter"][0] += nb
            tot["scatter"][1] += ms
            del gop, gx, goff
        out = torch.empty((B, h, w, O), device=xin.device, dtype=xin.dtype)
        w_conv = pr.w_off_tc if s == 1 else pr.w_off_s2d
        if w_conv is not None and L.ldconv_onepass_supported(B, C, H, W, N, s, O, O, dt):
            ms = timed(lambda: _lib.check(L.ldconv_onepass_fwd(xh.data_ptr(), w_conv.data_ptr(), pr.b_off.data_ptr(), pr.pn.data_ptr(),
                                                               pr.wt.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(), O, None,
                                                               B, C, H, W, N, s, O, _lib.ACT_SILU, dt, sv), "ldconv_onepass_fwd"))
            nb = x_bytes + out_bytes
            row.update({"onepass_MB": round(nb / 1e6, 1), "onepass_us": round(ms * 1e3, 1), "onepass_GBps": round(nb / ms / 1e6, 1),
                        "onepass_TFLOPs": round(2.0 * B * h * w * (N * C * O + 9 * C * 2 * N) / ms / 1e9, 1)})
            tot["one"][0] += nb
            tot["one"][1] += ms
            # the round-1 path on the same input: offset conv, then gather + GEMM
            t_off = timed(lambda: offset_conv_nhwc(xh, pr, N, s))
            t_gg = timed(lambda: _lib.check(L.ldconv_gather_gemm_fwd(xh.data_ptr(), off.data_ptr(), pr.pn.data_ptr(), pr.wt.data_ptr(),
                                                                     scale.data_ptr(), shift.data_ptr(), out.data_ptr(), O, B, C, H, W, N,
                                                                     s, O, _lib.ACT_SILU, dt, sv), "ldconv_gather_gemm_fwd"))
            row.update({"offconv_us": round(t_off * 1e3, 1), "gg_us": round(t_gg * 1e3, 1),
                        "gg_GBps": round((x_bytes + off_bytes + out_bytes) / t_gg / 1e6, 1)})
            tot["two"][0] += 2 * x_bytes + 2 * off_bytes + out_bytes
            tot["two"][1] += t_off + t_gg
        per_layer.append(row)
    peak = peaks.get("hbm_gbs", 6650.0)
    src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 (of fallback)"
    traffic, traffic_note = committed_traffic()
    big = max((r for r in per_layer if "onepass_us" in r), key=lambda r: r["onepass_MB"], default=None)
    roof = None
    if big is not None:
        ach = big["onepass_MB"] * 1e3 / big["onepass_us"]
        roof = {"bound": "hbm", "kernel": f"ldconv_onepass_kernel (whole LDConv forward: offset conv + sampling grid + gather + GEMM + BN + "
                f"SiLU, x read once; layer {big['layer']}: the largest LDConv launch of the step)", "achieved": round(ach, 1),
                "peak": peak, "peak_source": src, "unit": "GB/s", "frac": round(ach / peak, 4), "frac_of_8TBs_nominal": round(ach / 8000.0, 4),
                "traffic": traffic.get("roofline_kernel_bytes"), "traffic_source": traffic_note,
                "algorithmic_bytes_per_launch": round(big["onepass_MB"] * 1e6),
                "algorithmic_bytes": "e*B*C*H*W (x once) + e*B*h*w*O (out); offsets and the resampled operand stay on chip",
                "us_per_launch": big["onepass_us"], "timing": "CUDA events on the launching stream, median of %d, L2 flushed (256 MB "
                "memset) before every timed launch" % iters,
                "all_onepass_launches": {"GBps": round(tot["one"][0] / max(tot["one"][1], 1e-9) / 1e6, 1),
                                         "us_per_step": round(tot["one"][1] * 1e3, 1)},
                "round1_two_kernel_path": {"us_per_step": round(tot["two"][1] * 1e3, 1),
                                           "what": "offset conv (tcgen05) + ldconv_gather_gemm_fwd on the same inputs, for comparison"},
                "per_layer": per_layer}
    g_ach = tot["gather"][0] / max(tot["gather"][1], 1e-9) / 1e6
    roof_gather = {"bound": "hbm", "kernel": "gather_fwd_tiled_kernel (stand-alone LDConv resampling, training path; the 9 launches "
                   "with C >= 16 of one step)", "achieved": round(g_ach, 1), "peak": peak, "peak_source": src, "unit": "GB/s",
                   "frac": round(g_ach / peak, 4), "frac_of_8TBs_nominal": round(g_ach / 8000.0, 4),
                   "traffic": traffic.get("gather_kernel_bytes"),
                   "us_per_step": round(tot["gather"][1] * 1e3, 1)}
    roof_scatter = None
    if tot["scatter"][1] > 0:
        s_ach = tot["scatter"][0] / tot["scatter"][1] / 1e6
        big_s = max((r for r in per_layer if "scatter_us" in r), key=lambda r: r["scatter_MB"])
        roof_scatter = {"bound": "hbm", "kernel": "scatter_bwd_tiled2_kernel / gather_bwd_kernel (LDConv backward scatter of grad_x and "
                        "grad_offset, bf16 grad_x accumulator: ldconv_gather_bwd_acc16; the 9 launches with C >= 16 of one training step "
                        "at this batch)", "achieved": round(s_ach, 1), "peak": peak, "peak_source": src, "unit": "GB/s",
                        "frac": round(s_ach / peak, 4), "frac_of_8TBs_nominal": round(s_ach / 8000.0, 4), "traffic": None,
                        "algorithmic_bytes": "e*M*K (grad_operand) + e*B*C*H*W (x) + 4*B*2N*h*w (offsets) + e*B*C*H*W (grad_x) + 4*B*2N*h*w "
                        "(grad_offset)", "us_per_step": round(tot["scatter"][1] * 1e3, 1),
                        "largest_launch": {"layer": big_s["layer"], "us": big_s["scatter_us"], "GBps": big_s["scatter_GBps"],
                                           "frac": round(big_s["scatter_GBps"] / peak, 4)}}
    return roof, roof_gather, roof_scatter


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from experiment_yolo_b200 import _lib, dealyolo, engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: experiment_yolo_b200 has no CPU fallback "
                           "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if "BENCH_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["BENCH_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().ldconv_device_check(), "ldconv_device_check")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    torch.backends.cudnn.benchmark = True
    model = dealyolo.DealYolo(nc=NC)
    model.load_state_dict(dealyolo.seeded_state(model, 0))
    model = dealyolo.channels_last_(model.to(dev).bfloat16().eval())
    # the public inference call: the fused executor (every Conv / C2f / SPPF / ScalSeq / Detect block and every LDConv through
    # the library's kernels); --engine eager runs the plain torch graph around the CUDA LDConv instead
    if args.micro_batch:
        engine.FusedDealYolo.micro_batch = args.micro_batch
    if os.environ.get("BENCH_SERIAL_HEAD") == "1":       # A/B: Detect branches back to back on one stream
        engine._Detect.parallel_branches = False
    run = engine.FusedDealYolo(model) if args.engine == "fused" else model

    B = per_gpu_batch(world, args.scaling)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    n_in = n_input_buffers(B)      # the rotating set of input batches exceeds the 126 MB L2 at every per-GPU batch
    xs = [torch.rand((B, 3, IMG, IMG), device=dev, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
          for _ in range(n_in)]

    # ---- device-resident throughput: CUDA-graph replay of the whole forward ----------------------------------------------
    static_x = xs[0].clone()
    with torch.inference_mode():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                run(static_x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        _lib.call_counts.clear()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_y, _ = run(static_x)
        launches_per_step = sum(_lib.call_counts.values())

        def step(i):
            static_x.copy_(xs[i % n_in], non_blocking=True)
            graph.replay()

        for i in range(max(3, args.warmup)):
            step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record()
            for i in range(args.steps):
                step(i)
            e1.record()
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        ms_per_step = ms_max / args.steps
        value = B * world * args.steps / (ms_max / 1e3)

        # ---- end to end through the public call, host buffers in the timed region --------------------------------------
        # engine.PipelinedPredictor.submit()/result(): every step uploads its uint8 batch from pinned host memory, runs the
        # captured forward (uint8 -> bf16 NHWC conversion included) and downloads the detections to pinned host memory;
        # uploads / downloads of neighbouring steps overlap the compute on side streams, every result is waited for.
        host_u8 = [torch.randint(0, 256, (B, 3, IMG, IMG), dtype=torch.uint8).pin_memory() for _ in range(2)]

        def timed_e2e(e2e_run, drain=None):
            e2e_run(max(3, args.warmup))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record()
            e2e_run(args.steps)                      # returns only after the last result is in host memory
            if drain is not None:
                drain()
            e3.record()
            torch.cuda.synchronize()
            t2 = torch.tensor([e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            return B * world * args.steps / (float(t2.item()) / 1e3)

        e2e_extra = {}
        if args.engine == "fused":
            def make_run(pred):
                def e2e_run(n):
                    for i in range(n):
                        pred.submit(host_u8[i & 1])
                        if i >= 1:
                            pred.result()
                    pred.result()
                return e2e_run

            # (1) what the reference's predictor returns: detections after non_max_suppression (utils/ops.py:292, device-side
            # here).  The weights are random, so the class scores carry no objects: the confidence threshold is set once,
            # untimed, to the 99th percentile of the best-class scores of one batch (~336 candidates per image, a realistic load
            # for UAV scenes; at the default 0.25 every anchor of a random-init head is a candidate).
            probe = (host_u8[0].to(dev).float() / 255.0).bfloat16().contiguous(memory_format=torch.channels_last)
            y_probe, _ = run(probe)
            conf_thr = float(y_probe[:, 4:].float().amax(1).flatten().quantile(0.99))
            nms_kw = dict(conf_thres=conf_thr, iou_thres=0.45, max_det=300)
            pred = engine.PipelinedPredictor(model, B, IMG, nms=nms_kw)
            h2d, d2h = pred.h2d_bytes, pred.d2h_bytes
            e2e_value = timed_e2e(make_run(pred), pred.drain_to)
            det, cnt = pred.y_host[0]
            e2e_extra["nms"] = {"conf_thres": round(conf_thr, 5), "iou_thres": 0.45, "max_det": 300,
                                "kept_per_image_mean": round(float(cnt.float().mean()), 1),
                                "note": "conf_thres = 99th percentile of the best-class score (random-init weights), chosen untimed"}
            del pred
            # (2) the round-1 variant for comparison: the raw decoded head output (B, 4+nc, anchors) copied to the host
            pred_raw = engine.PipelinedPredictor(model, B, IMG)
            e2e_extra["raw_head_output"] = {"value": round(timed_e2e(make_run(pred_raw), pred_raw.drain_to), 2), "unit": "images/s",
                                            "d2h_bytes_per_step": pred_raw.d2h_bytes}
            del pred_raw
        else:
            host_out = torch.empty(tuple(static_y.shape), dtype=static_y.dtype).pin_memory()
            h2d, d2h = host_u8[0].numel(), host_out.numel() * host_out.element_size()

            def e2e_run(n):
                for i in range(n):
                    xu = host_u8[i & 1].to(dev, non_blocking=True)
                    xb = (xu.to(torch.bfloat16) * (1.0 / 255.0)).contiguous(memory_format=torch.channels_last)
                    y, _ = run(xb)
                    host_out.copy_(y, non_blocking=True)
                torch.cuda.synchronize()
            e2e_value = timed_e2e(e2e_run)

    # ---- BASELINE config 4 beside it: the training step (batch 128 split over the ranks, NCCL gradient all-reduce), so that the
    # driver's 1 / 2 / 4 / 8-GPU runs of this file also give the training curve.  Every rank takes part; a failure is reported,
    # it does not void the inference line.
    train = None
    if args.train_steps > 0 and args.scaling == "weak":
        try:
            del graph, static_y, static_x, xs
            torch.cuda.empty_cache()
            from benchmarks.train_step import run_train
            train = run_train(128, args.train_steps, 3, rank, world, dev)
        except Exception as e:
            train = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        torch.cuda.empty_cache()

    line = None
    if rank == 0:
        roof, roof_gather, roof_scatter = ldconv_roofline(model, host_u8[0].to(dev).float().div_(255.0).bfloat16().contiguous(memory_format=torch.channels_last),
                                            peaks, iters=7)
        cpu_best, cpu_mean, cpu_sec, cores = cpu_port_images_per_s(CPU_SAMPLE_BATCH, 3, 1) if world == 1 else (None,) * 4
        eager = gpu_eager_images_per_s(local, B) if world == 1 else None
        traffic, traffic_note = committed_traffic()
        step_bytes = traffic.get("step_dram_bytes")
        peak = peaks.get("hbm_gbs", 6650.0)
        line = {"metric": "images_per_sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 4),
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(world, args.engine, args.scaling), "lib_sha16": lib_sha16(),
                "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "api": "engine.PipelinedPredictor.submit()/result()" if args.engine == "fused"
                        else "DealYolo.forward", "input": "uint8 NCHW batch in pinned host memory, normalised on device",
                        "result": "detections after device-side non_max_suppression, (B,300,6) fp32 + counts, copied to pinned host "
                                  "memory" if args.engine == "fused" else "decoded head output (B,10,33600) bf16 copied to pinned host memory",
                        **e2e_extra},
                "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
                "clocks": clk.summary(), "roofline": roof, "roofline_gather": roof_gather, "roofline_scatter": roof_scatter}
        if step_bytes and B == PER_GPU_BATCH:
            # whole-step roofline: DRAM bytes of one step (sum over the ncu launch list of this library build) / step time
            line["step_roofline"] = {"bound": "hbm", "dram_bytes_per_step": step_bytes, "achieved": round(step_bytes / ms_per_step / 1e6, 1),
                                     "peak": peak, "unit": "GB/s", "frac": round(step_bytes / ms_per_step / 1e6 / peak, 4),
                                     "source": traffic_note}
        if train is not None:
            line["config4_train"] = train
        if eager is not None:
            line["gpu_eager_baseline"] = eager
        if cpu_mean is not None:
            line["cpu_baseline"] = {"value": round(cpu_mean, 3), "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"batch {CPU_SAMPLE_BATCH} x 3 forwards of the same graph in fp32 (eager CPU port "
                                              f"of the reference LDConv, best {cpu_best:.2f} images/s)"}
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else a library prints lands on stderr (see main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    # keep stdout to the one JSON line: NCCL prints its version banner to fd 1, torch / ctypes libraries may print too
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default="fused", choices=["fused", "eager"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 64 images per GPU; strong: one 64-image batch split over the ranks (SURVEY.md 8d config 3)")
    ap.add_argument("--micro-batch", type=int, default=0, help="images per pass of the fused executor (0 = its default)")
    ap.add_argument("--train-steps", type=int, default=6, help="timed steps of the config-4 training leg (0 = skip)")
    ap.add_argument("--cpu-batch", type=int, default=0, help="--impl reference: images per timed step (default 8; 1 = BASELINE config 1)")
    ap.add_argument("--eager-leg", default=None, help=argparse.SUPPRESS)      # internal: one dtype of gpu_eager_baseline
    ap.add_argument("--eager-batch", type=int, default=PER_GPU_BATCH, help=argparse.SUPPRESS)
    ap.add_argument("--eager-iters", type=int, default=3, help=argparse.SUPPRESS)
    ap.add_argument("--eager-device", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.eager_leg:
        return gpu_eager_leg(args.eager_leg, args.eager_batch, args.eager_iters, args.eager_device)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
