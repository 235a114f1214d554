"""Developer timing probe for config 4: N synthetic 3x32x32 images against a random-init torchvision ResNet-18
(num_classes=10, eval, fp32), 8x8 grid of 4x4 segments, K ADMM iterations of update_G; compared with the reference's own
formulation (oracle/sa_oracle.py == main_ori.py:626-743: batch 1, PyTorch ops, two .item() syncs per iteration)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torchvision
from lpbox import sparse_attack as sa
from sa_util import grid_segments
import numpy as np
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
torch.manual_seed(0)
model = torchvision.models.resnet18(num_classes=10).cuda().eval()
for p in model.parameters(): p.requires_grad_(False)
images = torch.rand(N, 3, 32, 32, device="cuda")
with torch.no_grad(): target = (model(images - 0.5).argmax(1) + 1) % 10
eps = 0.1 * torch.randn(N, 3, 32, 32, device="cuda")
G0 = torch.ones(N, 3, 32, 32, device="cuda"); nw = torch.ones_like(G0)
seg = torch.from_numpy(np.broadcast_to(grid_segments(), (3, 32, 32)).copy()).reshape(-1).to(torch.int32).cuda()
sa.update_G(model, images, target, eps, G0.clone(), sa.init_params(), seg, nw, args={"maxIter_g": 3})
torch.cuda.synchronize(); t = time.time()
G, res = sa.update_G(model, images, target, eps, G0.clone(), sa.init_params(), seg, nw, args={"maxIter_g": K})
torch.cuda.synchronize(); dt = time.time() - t
print(f"ours: N={N} K={K}: {dt:.3f}s  {N*K/dt:.3e} image-iterations/s  ({N/dt*K/2000:.1f} images/s at 2000 iterations)  G.sum mean {float(G.sum((1,2,3)).mean()):.2f}")
import sa_oracle
B = torch.zeros(64, 3, 32, 32, device="cuda")
sg = torch.from_numpy(grid_segments()).cuda()
for s in range(64): B[s, :, sg == s] = 1
mean = torch.full((1, 3, 1, 1), 0.5, device="cuda"); std = torch.ones((1, 3, 1, 1), device="cuda")
Kr = min(K, 50)
sa_oracle.update_G(model, images[:1], target[:1], eps[:1], G0[:1].clone(), sa_oracle.INIT, B, nw[:1], 3, mean=mean, std=std)
torch.cuda.synchronize(); t = time.time()
Go, _, _ = sa_oracle.update_G(model, images[:1], target[:1], eps[:1], G0[:1].clone(), sa_oracle.INIT, B, nw[:1], Kr, mean=mean, std=std)
torch.cuda.synchronize(); dr = time.time() - t
print(f"reference formulation on the same GPU (batch 1): {Kr/dr:.1f} image-iterations/s; ratio {N*K/dt/(Kr/dr):.0f}x; max|dG| vs ours after {Kr} its: {float((G[:1]-Go).abs().max()) if Kr==K else float('nan'):.2e}")
