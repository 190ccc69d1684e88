#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py -x -q -m gpu -k "gather_gemm or fused or model or module" > gpurun_out/s4_gg_tests.log 2>&1; echo "gg tests exit $?"; tail -3 gpurun_out/s4_gg_tests.log
LDCONV_GG_SPLIT=0 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s4_nosplit.json 2> gpurun_out/bench_s4_nosplit.err; echo "bench exit $?"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s4f.json 2> gpurun_out/bench_s4f.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("bench_s4_nosplit","bench_s4f"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, d["value"], d["e2e"]["value"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["roofline"]["all_gg_launches"], [(r["layer"], r.get("gg_us")) for r in d["roofline"]["per_layer"]])
PY
