#!/usr/bin/env python
"""bench.py -- DEAL-YOLO-LD 640x640 images/s on B200 (BASELINE.json metric), with the LDConv roofline beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one full forward of DEAL-YOLO-LD (cfg/deal-yolo-ld-p2.yaml = the reference's yolov8-LD-P2.yaml, 918,304
parameters, random-init seeded weights) over one synthetic batch of 64 bf16 640x640 images per GPU, inference mode,
channels_last; the batch is sharded over ranks with no collective (every LDConv sample depends on its own image only).
  value     images/s with the batch already resident in HBM (CUDA-graph replay of the forward), max over ranks
  e2e       images/s through the public call `engine.FusedDealYolo(model)(images)` with HOST buffers: every step copies its uint8 batch from
            pinned host memory, normalises on the device, runs the forward and reads the detections back to the host
  roofline  the LDConv gather kernels of one step (10 launches): algorithmic bytes / CUDA-event time
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
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
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
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
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
                "samples": len(sm)}


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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    best, mean, sec, cores = cpu_port_images_per_s(CPU_SAMPLE_BATCH, max(1, args.steps), max(1, min(args.warmup, 2)))
    sample = f"batch {CPU_SAMPLE_BATCH} of 640x640 fp32 images per step, eager CPU port of the reference graph"
    line = {"impl": "reference", "metric": "images_per_sec", "value": round(mean, 3), "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": round(mean, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(mean, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus: int, engine: str = "fused"):
    return {"engine": engine, "workload": "DEAL-YOLO-LD (yolov8-LD-P2 graph, 10 LDConv + SSFF, nc=6) full forward, 640x640, inference",
            "global_batch": PER_GPU_BATCH * n_gpus, "per_gpu_batch": PER_GPU_BATCH, "imgsz": IMG,
            "layout": "channels_last", "parallelism": f"batch-sharded x{n_gpus}, no collective",
            "l2_policy": "inputs_exceed_l2 (157 MB bf16 batch, two alternating input buffers; activations of one step are GBs)",
            "weights": "seeded random init (dealyolo.seeded_state(0)), p_conv.weight ~ N(0,0.05)"}


# --------------------------------------------------------------------------------------------------------------- GPU arm
def ldconv_gather_roofline(model, x, peaks, iters: int):
    """Time the gather kernel of each of the 10 LDConv layers at this step's shapes with CUDA events on the launching
    stream, feeding it the real layer inputs; achieved = algorithmic bytes / time (SURVEY.md 8d formula)."""
    import torch
    from experiment_yolo_b200 import _lib
    L = _lib.load()
    feats = {}
    hooks = [m.register_forward_pre_hook(lambda mod, inp, i=m.i: feats.__setitem__(i, inp[0])) for m in model.ldconv_layers()]
    with torch.inference_mode():
        model(x)
    for h in hooks:
        h.remove()
    st = torch.cuda.current_stream()
    tot_bytes, tot_ms, per_layer = 0.0, 0.0, []
    for m in model.ldconv_layers():
        xin = feats[m.i]
        B, C, H, W = xin.shape
        N, s = m.num_param, int(m.stride)
        h, w = (H - 1) // s + 1, (W - 1) // s + 1
        e = xin.element_size()
        xh = xin.permute(0, 2, 3, 1).contiguous()
        pr = m._prepared(xin.dtype, False)
        off = torch.empty((B, h, w, 2 * N), device=xin.device, dtype=torch.float32)
        dt = _lib.BF16 if xin.dtype == torch.bfloat16 else _lib.F32
        _lib.check(L.ldconv_offset_conv_fwd(xh.data_ptr(), pr.w_off.data_ptr(), pr.b_off.data_ptr(), off.data_ptr(), B, C, H,
                                            W, N, s, dt, st.cuda_stream), "ldconv_offset_conv_fwd")
        operand = torch.empty((B * h * w, N * C), device=xin.device, dtype=xin.dtype)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in ev:
            a.record(st)
            _lib.check(L.ldconv_gather_fwd(xh.data_ptr(), off.data_ptr(), pr.pn.data_ptr(), operand.data_ptr(), None, None,
                                           B, C, H, W, N, s, dt, st.cuda_stream), "ldconv_gather_fwd")
            b.record(st)
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        ms = sum(ms[: max(1, len(ms) // 2 + 1)]) / max(1, len(ms) // 2 + 1) if len(ms) > 2 else sum(ms) / len(ms)
        nbytes = e * B * C * H * W + 4 * B * 2 * N * h * w + e * B * h * w * N * C
        tot_bytes += nbytes
        tot_ms += ms
        per_layer.append({"layer": m.i, "C": C, "N": N, "s": s, "hw": [h, w], "MB": round(nbytes / 1e6, 1),
                          "us": round(ms * 1e3, 1), "GBps": round(nbytes / ms / 1e6, 1)})
    achieved = tot_bytes / tot_ms / 1e6
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"bound": "hbm", "kernel": "ldconv gather_fwd (10 launches of one step)", "achieved": round(achieved, 1),
            "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else
            "B200_PROFILING.md fallback 6650 (of fallback)", "unit": "GB/s", "frac": round(achieved / peak, 4),
            "frac_of_8TBs_nominal": round(achieved / 8000.0, 4), "traffic": None,
            "algorithmic_bytes_per_launch": round(tot_bytes / len(per_layer)), "avg_us_per_launch":
            round(tot_ms * 1e3 / len(per_layer), 2), "per_layer": per_layer}


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
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line
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
    run = engine.FusedDealYolo(model) if args.engine == "fused" else model

    B = PER_GPU_BATCH
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    xs = [torch.rand((B, 3, IMG, IMG), device=dev, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
          for _ in range(2)]

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
            static_x.copy_(xs[i & 1], non_blocking=True)
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
        if args.engine == "fused":
            pred = engine.PipelinedPredictor(model, B, IMG)
            h2d, d2h = pred.h2d_bytes, pred.d2h_bytes

            def e2e_run(n):
                for i in range(n):
                    pred.submit(host_u8[i & 1])
                    if i >= 1:
                        pred.result()
                pred.result()
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

        e2e_run(max(3, args.warmup))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        e2e_run(args.steps)                      # returns only after the last result is in host memory
        if args.engine == "fused":
            pred.drain_to()
        e3.record()
        torch.cuda.synchronize()
        t2 = torch.tensor([e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_value = B * world * args.steps / (float(t2.item()) / 1e3)

    line = None
    if rank == 0:
        roof = ldconv_gather_roofline(model, xs[0], peaks, iters=5)
        cpu_best, cpu_mean, cpu_sec, cores = cpu_port_images_per_s(CPU_SAMPLE_BATCH, 3, 1) if world == 1 else (None,) * 4
        line = {"metric": "images_per_sec", "value": round(value, 2), "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 4),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(world, args.engine),
                "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "api": "engine.PipelinedPredictor.submit()/result()" if args.engine == "fused"
                        else "DealYolo.forward", "input": "uint8 NCHW batch in pinned host memory, normalised on device",
                        "result": "decoded detections (B,10,33600) bf16 copied to pinned host memory"},
                "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
                "clocks": clk.summary(), "roofline": roof}
        if cpu_mean is not None:
            line["cpu_baseline"] = {"value": round(cpu_mean, 3), "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"batch {CPU_SAMPLE_BATCH} x 3 forwards of the same graph in fp32 (eager CPU port "
                                              f"of the reference LDConv, best {cpu_best:.2f} images/s)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default="fused", choices=["fused", "eager"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
