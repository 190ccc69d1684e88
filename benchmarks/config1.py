"""BASELINE.json config 1: DEAL-YOLO-LD forward, batch 1, synthetic 640 x 640 -- latency on the GPU next to the CPU path.
  gpu fp32 module graph : dealyolo.DealYolo (CUDA LDConv, torch / cuDNN around it), fp32, the precision the parity fixture is held to
  gpu bf16 fused engine : engine.FusedDealYolo, CUDA-graph replay (what bench.py runs at batch 64)
  cpu port              : `bench.py --impl reference --cpu-batch 1` (the eager CPU port of the reference graph on all host threads; the
                          reference itself cannot travel to the GPU box, oracle/gen_model_golden.py measured port vs reference where both
                          exist; bench.py's reference arm is the one place outside tests/ that may execute oracle/)
    python benchmarks/config1.py [--iters 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import dealyolo, engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu-iters", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    x = torch.rand((1, 3, 640, 640))

    def gpu_ms(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    model = dealyolo.DealYolo(nc=6)
    model.load_state_dict(dealyolo.seeded_state(model, 0))
    m32 = dealyolo.channels_last_(model.to(dev).eval())
    x32 = x.to(dev).contiguous(memory_format=torch.channels_last)
    with torch.inference_mode():
        med, best = gpu_ms(lambda: m32(x32))
    print(json.dumps({"config": 1, "path": "gpu fp32 module graph (CUDA LDConv + torch/cuDNN)", "batch": 1, "imgsz": 640, "ms_median": round(med, 3),
                      "ms_min": round(best, 3), "images_per_s": round(1e3 / med, 1)}), flush=True)
    mb = dealyolo.channels_last_(model.bfloat16().eval())
    run = engine.FusedDealYolo(mb)
    xb = x.to(dev).bfloat16().contiguous(memory_format=torch.channels_last)
    with torch.inference_mode():
        for _ in range(3):
            run(xb)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run(xb)
        med, best = gpu_ms(g.replay)
    print(json.dumps({"config": 1, "path": "gpu bf16 fused engine, CUDA-graph replay", "batch": 1, "imgsz": 640, "ms_median": round(med, 3),
                      "ms_min": round(best, 3), "images_per_s": round(1e3 / med, 1)}), flush=True)
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--cpu-batch", "1", "--steps", str(args.cpu_iters),
                          "--warmup", "1"], capture_output=True, text=True, cwd=root)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    print(json.dumps({"config": 1, "path": "cpu eager port of the reference graph, fp32 (bench.py --impl reference --cpu-batch 1)", "batch": 1,
                      "imgsz": 640, "threads": line["cpu_baseline"]["cores"], "ms_median": line["ms_per_step"],
                      "images_per_s": line["value"]}), flush=True)

if __name__ == "__main__":
    main()
