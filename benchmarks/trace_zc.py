"""Timeline (clock64) of the zero-copy 3x3 conv kernel's three roles for tile iterations 8..11 of CTA 0 (LDCONV_DBG=32).
    python benchmarks/trace_zc.py --cin 64 --cout 64 --hw 160
"""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["LDCONV_DBG"] = "32"
from experiment_yolo_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=64); ap.add_argument("--cout", type=int, default=64)
ap.add_argument("--hw", type=int, default=160); ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
L = _lib.load()
dev = torch.device("cuda", 0)
B, H = a.batch, a.hw
x = torch.randn((B, H, H, a.cin), device=dev).bfloat16()
w = (torch.randn((a.cout, 9 * a.cin), device=dev) * 0.05).bfloat16()
sc, sh = torch.ones(a.cout, device=dev), torch.zeros(a.cout, device=dev)
out = torch.empty((B, H, H, a.cout), device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
fn = ctypes.CDLL(_lib.LIB_PATH).ldconv_debug_trace_zc
buf = (ctypes.c_longlong * 8192)()
for rep in range(2):
    _lib.check(L.ldconv_conv3x3_bn_act_fwd(x.data_ptr(), a.cin, w.data_ptr(), sc.data_ptr(), sh.data_ptr(), None, 0, out.data_ptr(),
                                           a.cout, B, a.cin, H, H, a.cout, 1, _lib.ACT_SILU, _lib.BF16, st))
    torch.cuda.synchronize()
    n = fn(buf, 4096)
ev = sorted(((buf[2 * i + 1], buf[2 * i]) for i in range(n)))
t0 = ev[0][0]
names = {1: "TMA issue", 2: "MMA operands ready", 3: "MMA issued+committed", 4: "EPI wait t_full", 5: "EPI got t_full",
         6: "MMA wait t_empty", 7: "MMA got t_empty", 9: "EPI done"}
for t, tag in ev:
    role, rest = divmod(tag, 100000)
    it, kb = divmod(rest, 100)
    print(f"{t - t0:8d}  it={it:2d}  {names.get(role, role)}")
