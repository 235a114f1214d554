"""Developer timing probe of the fused feed-forward kernel alone: M rows x (128 -> 512 -> 128)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import torch
import lpbox
L = lpbox._capi.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 655360
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
X = (torch.randn(M, 128, device="cuda") * 0.5).bfloat16()
W1 = (torch.randn(512, 128, device="cuda") * 0.1).bfloat16(); W2 = (torch.randn(128, 512, device="cuda") * 0.1).bfloat16()
b1 = torch.randn(512, device="cuda"); b2 = torch.randn(128, device="cuda"); sc = torch.rand(128, device="cuda") + 0.5; sh = torch.randn(128, device="cuda")
out = torch.zeros(M, 128, device="cuda", dtype=torch.bfloat16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = lambda t: C.c_void_p(t.data_ptr())
run = lambda: L.lpbox_ff_fused_dev(st, vp(X), vp(W1), vp(b1), vp(W2), vp(b2), vp(sc), vp(sh), vp(out), M)
for _ in range(3): assert run() == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"M={M}: {us:.1f} us  {4 * 128 * 512 * M / us / 1e6:.1f} TFLOP/s  {3 * M * 256 / us / 1e3:.0f} GB/s (X read twice + out)")
