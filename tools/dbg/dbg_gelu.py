import torch, torch.nn.functional as F, sys
sys.path.insert(0,'/root/repo')
from graph_augmented_vision_transformers_b200 import ops
DEV='cuda'
g = torch.Generator().manual_seed(0)
u = (torch.randn(64, 197, 3072, generator=g) * 1.5)
cot = torch.randn(64, 197, 3072, generator=g)
ud = u.to(DEV).requires_grad_(True)
torch.manual_seed(1)
out = ops.gelu_dropout(ud, 0.1, training=True)
out.backward(cot.to(DEV))
ref = F.gelu(u).to(DEV)/0.9
keptg = ud.grad != 0
kepto = out != 0
print('kept frac grad', keptg.float().mean().item(), 'out', kepto.float().mean().item())
mism = keptg != kepto
print('mismatch count', mism.sum().item())
idx = mism.nonzero()[:10]
for i in idx:
    i=tuple(i.tolist()); print(i, 'u', u[i].item(), 'out', out[i].item(), 'grad', ud.grad[i].item(), 'ref', ref[i].item())
err = ((out-ref).abs()*keptg)
print('max err on kept', err.max().item(), 'at u', u.to(DEV).flatten()[err.argmax()].item())
