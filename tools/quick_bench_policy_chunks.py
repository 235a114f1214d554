import os, sys
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import torch
from lpbox.policy import GraphAttentionEncoder
from lpbox.policy_kernel import PolicyKernel
rows = 500000
torch.manual_seed(0)
net = GraphAttentionEncoder(tokens=20).cuda().eval()
x = torch.rand(rows, 20, 5, device="cuda")
macs = 20 * 2 * (128 * 384 + 128 * 128 + 128 * 512 * 2) + 2560 * 256 + 256 * 128 + 128 * 16 + 16 + 20 * 10 * 128 + 2 * 2 * 8 * 20 * 20 * 16
for chunk in (16384, 32768, 65536, 131072):
    pk = PolicyKernel(net, chunk_rows=chunk)
    for _ in range(2): pk(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = pk(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"chunk={chunk} {ms:.2f} ms  {2*macs*rows/ms/1e9:.1f} TFLOP/s")
    pk.close()
