import ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["LDCONV_DBG"] = "64"
from benchmarks.ldconv_layers import LAYERS
from experiment_yolo_b200 import _lib
L = _lib.load()
dev = torch.device("cuda", 0)
fn = ctypes.CDLL(_lib.LIB_PATH).ldconv_debug_fused_trace
names = ["start", "barriers+tmem alloc", "TMA issued + w_off staged", "x tile landed", "offset conv done", "gather done (+fence,sync)",
         "MMA issued+committed", "MMA complete", "epilogue done"]
for (li, C, O, N, s, H) in LAYERS[1:]:
    B, W = 64, H
    h = w = (H - 1) // s + 1
    x = torch.randn((B, H, W, C), device=dev).bfloat16()
    w_off = torch.randn((3, 3, C, 2 * N), device=dev) * 0.05
    b_off = torch.zeros(2 * N, device=dev)
    pn = torch.tensor(_lib.p_n_table(N), dtype=torch.int32, device=dev)
    wt = (torch.randn((O, N * C), device=dev) * 0.1).bfloat16()
    scale, shift = torch.ones(O, device=dev), torch.zeros(O, device=dev)
    out = torch.empty((B * h * w, O), device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    if not L.ldconv_fused_supported(B, C, H, W, N, s, O, 1):
        continue
    for _ in range(2):
        _lib.check(L.ldconv_fused_fwd(x.data_ptr(), w_off.data_ptr(), b_off.data_ptr(), pn.data_ptr(), wt.data_ptr(), scale.data_ptr(),
                                      shift.data_ptr(), out.data_ptr(), None, B, C, H, W, N, s, O, 1, 1, st))
        torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    fn(buf)
    t = [buf[i] - buf[0] for i in range(9)]
    print(f"L{li} C{C} N{N} s{s} O{O}: " + ", ".join(f"{names[i]}={t[i]}" for i in range(1, 9)))
