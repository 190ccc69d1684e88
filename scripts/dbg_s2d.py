import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from experiment_yolo_b200 import _lib
from experiment_yolo_b200.ldconv import _prepare, base_grid
L = _lib.load(); dev = 'cuda:0'
B, C, H, W, N = 1, 16, 8, 16, 1
x = torch.zeros(B, C, H, W)
for r in range(H):
    for k in range(W):
        x[0, :, r, k] = r * 16 + k + torch.arange(C) * 0   # position code (exact in bf16 up to 256)
x = x.bfloat16()
for (ky, kx, c0) in [(1, 1, 0), (0, 1, 0), (2, 1, 0), (1, 0, 0), (1, 2, 0), (1, 1, 5), (2, 2, 9)]:
    w = torch.zeros(2 * N, C, 3, 3); w[0, c0, ky, kx] = 1
    pr = _prepare(w.bfloat16().to(dev), torch.zeros(2 * N, device=dev), torch.zeros(8, C, N, 1, device=dev), base_grid(N), torch.bfloat16, False)
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    off = torch.full((B, H // 2, W // 2, 2 * N), float('nan'), device=dev)
    _lib.check(L.ldconv_offset_conv_s2d_fwd(xd.data_ptr(), pr.w_off_s2d.data_ptr(), pr.b_off.data_ptr(), off.data_ptr(), B, C, H, W, N, _lib.BF16, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = off[0, :, :, 0].cpu().numpy()
    want = np.zeros((H // 2, W // 2))
    for i in range(H // 2):
        for j in range(W // 2):
            r, k = 2 * i + ky - 1, 2 * j + kx - 1
            want[i, j] = (r * 16 + k) if (0 <= r < H and 0 <= k < W) else 0
    print('tap', ky, kx, 'c', c0, 'max diff', np.abs(got - want).max())
    if np.abs(got - want).max() > 0:
        print(' got row0', got[0], '\n got row1', got[1]); print(' want row0', want[0], '\n want row1', want[1])

print('--- random case')
from oracle import oracle
for (C, N, H, W, B) in [(16, 1, 8, 16, 1), (16, 3, 8, 16, 1), (16, 3, 32, 16, 1), (16, 3, 8, 32, 1), (16, 3, 40, 56, 1), (16, 3, 40, 56, 2)]:
    g = torch.Generator().manual_seed(C + N + H)
    x = torch.randn(B, C, H, W, generator=g).bfloat16()
    w = (torch.randn(2 * N, C, 3, 3, generator=g) * 0.1).bfloat16()
    b = torch.randn(2 * N, generator=g)
    want = oracle.offset_conv(x.float().numpy(), w.float().numpy(), b.numpy(), N, 2)
    pr = _prepare(w.to(dev), b.to(dev), torch.zeros(8, C, N, 1, device=dev), base_grid(N), torch.bfloat16, False)
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    off = torch.full((B, H // 2, W // 2, 2 * N), float('nan'), device=dev)
    _lib.check(L.ldconv_offset_conv_s2d_fwd(xd.data_ptr(), pr.w_off_s2d.data_ptr(), pr.b_off.data_ptr(), off.data_ptr(), B, C, H, W, N, _lib.BF16, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = off.cpu().numpy().transpose(0, 3, 1, 2)
    d = np.abs(got - want)
    bad = np.argwhere(d > 1e-3)
    print((C, N, H, W, B), 'max diff', d.max(), 'nbad', len(bad), 'of', d.size, 'first bad', bad[:6].tolist())
