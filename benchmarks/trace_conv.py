import ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["LDCONV_DBG"] = "32"
from experiment_yolo_b200 import _lib
L = _lib.load()
dev = torch.device("cuda", 0)
B, C, H, W, N, s = 64, 32, 160, 160, 1, 1
x = torch.randn((B, H, W, C), device=dev).bfloat16()
w = (torch.randn((2 * N, 9 * C), device=dev) * 0.05).bfloat16()
b = torch.zeros(2 * N, device=dev)
off = torch.empty((B, H, W, 2 * N), device=dev)
st = torch.cuda.current_stream().cuda_stream
fn = ctypes.CDLL(_lib.LIB_PATH).ldconv_debug_trace_zc
buf = (ctypes.c_longlong * 8192)()
for rep in range(2):
    _lib.check(L.ldconv_offset_conv_tc_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), off.data_ptr(), B, C, H, W, N, s, 1, st))
    torch.cuda.synchronize()
    n = fn(buf, 4096)
ev = sorted(((buf[2 * i + 1], buf[2 * i]) for i in range(n)))
t0 = ev[0][0]
names = {1: "TMA issue", 2: "MMA operands ready", 3: "MMA issued+committed", 4: "EPI wait t_full", 5: "EPI got t_full",
         6: "MMA wait t_empty", 7: "MMA got t_empty", 8: "WRK copies done", 9: "EPI done"}
for t, tag in ev:
    role, rest = divmod(tag, 100000)
    it, kb = divmod(rest, 100)
    print(f"{t - t0:8d}  it={it:2d} kb={kb:2d}  {names.get(role, role)}")
