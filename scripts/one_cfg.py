import sys, os, torch
sys.path.insert(0, '/root/repo')
import experiment_yolo_b200 as E
from experiment_yolo_b200 import _lib
C, N, H, s, B = [int(v) for v in sys.argv[1:6]]
dev = torch.device('cuda', 0)
torch.manual_seed(0)
mod = E.LDConv(C, C, N, s).to(dev)
with torch.no_grad():
    mod.p_conv.weight.normal_(0, 0.05)
mod = mod.bfloat16().eval()
x = torch.randn(B, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
_lib.call_counts.clear()
with torch.no_grad():
    y = mod(x)
torch.cuda.synchronize()
print('ok', sorted(_lib.call_counts), float(y.float().abs().mean()))
