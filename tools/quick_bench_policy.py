"""Developer timing probe of the tensor-core policy: rows x (T=20, 5) -> scores; reports TFLOP/s (17.56 MFLOP per row)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import torch
from lpbox.policy import GraphAttentionEncoder
from lpbox.policy_kernel import PolicyKernel
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
torch.manual_seed(0)
net = GraphAttentionEncoder(tokens=20).cuda().eval()
pk = PolicyKernel(net, chunk_rows=32768)
x = torch.rand(rows, 20, 5, device="cuda")
for _ in range(2): pk(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = pk(x); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
macs = 20 * 2 * (128 * 384 + 128 * 128 + 128 * 512 * 2) + 2560 * 256 + 256 * 128 + 128 * 16 + 16 + 20 * 10 * 128 + 2 * 2 * 8 * 20 * 20 * 16
print(f"rows={rows} kernel path: {ms:.1f} ms  {rows/ms*1e3:.3e} rows/s  {2*macs*rows/ms/1e9:.1f} TFLOP/s")
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(2): [net(x[a:a+32768])[1] for a in range(0, rows, 32768)]
    torch.cuda.synchronize(); e0.record(); [net(x[a:a+32768])[1] for a in range(0, rows, 32768)]; e1.record(); torch.cuda.synchronize()
print(f"torch bf16 autocast: {e0.elapsed_time(e1):.1f} ms")
