import sys, torch
sys.path.insert(0, '/root/repo')
from experiment_yolo_b200 import dealyolo, _lib
from experiment_yolo_b200 import dist as xdist
import experiment_yolo_b200.ldconv as ld
dev = torch.device('cuda', 0)
model = dealyolo.DealYolo(nc=6)
model.load_state_dict(dealyolo.seeded_state(model, 0))
model = dealyolo.channels_last_(model.to(dev)).train()
B = 16
x = torch.rand((B, 3, 640, 640), device=dev).contiguous(memory_format=torch.channels_last)
targets = [torch.zeros((B, 70, 640 // s, 640 // s), device=dev) for s in (4, 8, 16)]
opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.01, momentum=0.937, nesterov=True)
stats = {}
def hook(mod, inp):
    xx = inp[0].detach().float()
    off = torch.nn.functional.conv2d(xx, mod.p_conv.weight.float(), mod.p_conv.bias.float(), stride=mod.stride, padding=1)
    stats[mod.i] = (float(off.abs().mean()), float(off.abs().max()), str(inp[0].dtype))
for m in model.ldconv_layers():
    m.register_forward_pre_hook(hook)
for it in range(5):
    with torch.autocast(device_type='cuda', dtype=torch.bfloat16):
        opt.zero_grad(set_to_none=True)
        outs = model(x)
    loss = xdist.surrogate_detection_loss(outs, targets)
    loss.backward()
    opt.step()
    print('step', it, 'loss', float(loss), {k: (round(v[0], 2), round(v[1], 1), v[2][6:]) for k, v in stats.items()})
