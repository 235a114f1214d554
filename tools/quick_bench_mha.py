"""Developer timing probe of the fused multi-head-attention sublayer kernel alone: R variables x T tokens x 128."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "accelerated-lpbox-admm_b200"))
import torch
import lpbox
L = lpbox._capi.lib()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
M = R * T
X = (torch.randn(M, 128, device="cuda") * 0.7).bfloat16()
Wqkv = (torch.randn(384, 128, device="cuda") * 0.12).bfloat16(); Wo = (torch.randn(128, 128, device="cuda") * 0.1).bfloat16()
sc = torch.rand(128, device="cuda") + 0.5; sh = torch.randn(128, device="cuda")
out = torch.zeros(M, 128, device="cuda", dtype=torch.bfloat16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = lambda t: C.c_void_p(t.data_ptr())
run = lambda: L.lpbox_mha_fused_dev(st, vp(X), vp(Wqkv), vp(Wo), vp(sc), vp(sh), vp(out), M, T)
for _ in range(3): assert run() == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
flop = 2.0 * M * (128 * 384 + 128 * 128) + 2.0 * R * 8 * (2 * T * T * 16)
print(f"R={R} T={T} M={M}: {us:.1f} us  {flop / us / 1e6:.1f} TFLOP/s  {2 * M * 256 / us / 1e3:.0f} GB/s (X + out)")
