"""GPU tests of the tensor-core policy path (D1): the tcgen05 bf16 GEMM against torch, and the whole network against the
fp32 PyTorch module (tolerance: bf16 activations/weights with fp32 accumulation -> |d sigmoid| <= 0.02)."""
import ctypes as C

import numpy as np
import pytest
import torch

from policy_weights import fill_deterministic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 384, 128), (1000, 128, 512), (77, 256, 2560), (4096, 512, 128)])
def test_tcgen05_gemm_matches_torch(M, N, K):
    import lpbox
    L = lpbox._capi.lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    Cc = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.lpbox_gemm_bf16_dev(st, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(Cc.data_ptr()), M, N, K, C.c_void_p(bias.data_ptr()), 1)
    assert rc == 0, lpbox._capi.last_error()
    torch.cuda.synchronize()
    ref = torch.relu(A.float() @ W.float().t() + bias)
    err = (Cc.float() - ref).abs().max().item()
    assert err <= 0.02 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M", [128, 256, 300, 5000, 100000])
def test_fused_feed_forward_matches_torch(M):
    import lpbox
    L = lpbox._capi.lib()
    g = torch.Generator(device="cuda").manual_seed(M)
    X = (torch.randn(M, 128, device="cuda", generator=g) * 0.5).bfloat16()
    W1 = (torch.randn(512, 128, device="cuda", generator=g) * 0.1).bfloat16()
    W2 = (torch.randn(128, 512, device="cuda", generator=g) * 0.1).bfloat16()
    b1 = torch.randn(512, device="cuda", generator=g) * 0.2
    b2 = torch.randn(128, device="cuda", generator=g) * 0.2
    sc = torch.rand(128, device="cuda", generator=g) + 0.5
    sh = torch.randn(128, device="cuda", generator=g) * 0.1
    out = torch.zeros(M, 128, device="cuda", dtype=torch.bfloat16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())
    rc = L.lpbox_ff_fused_dev(st, vp(X), vp(W1), vp(b1), vp(W2), vp(b2), vp(sc), vp(sh), vp(out), M)
    assert rc == 0, lpbox._capi.last_error()
    torch.cuda.synchronize()
    H = torch.relu(X.float() @ W1.float().t() + b1).bfloat16().float()      # the hidden activation is rounded to bf16 on chip
    ref = (X.float() + H @ W2.float().t() + b2) * sc + sh
    err = (out.float() - ref).abs().max().item()
    assert err <= 0.02 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("T,R", [(20, 6), (20, 7), (20, 1000), (10, 12), (10, 501), (5, 25), (5, 1003), (20, 40000)])
def test_fused_mha_sublayer_matches_torch(T, R):
    """out = (X + Wo MHA(X)) * scale + shift for R variables of T tokens vs an fp32 PyTorch restatement (q, k, v and the head
    outputs are rounded to bf16 on chip, like the unfused kernels)."""
    import lpbox
    L = lpbox._capi.lib()
    g = torch.Generator(device="cuda").manual_seed(T * 1000 + R)
    M = R * T
    X = (torch.randn(M, 128, device="cuda", generator=g) * 0.7).bfloat16()
    Wqkv = (torch.randn(384, 128, device="cuda", generator=g) * 0.12).bfloat16()
    Wo = (torch.randn(128, 128, device="cuda", generator=g) * 0.1).bfloat16()
    sc = torch.rand(128, device="cuda", generator=g) + 0.5
    sh = torch.randn(128, device="cuda", generator=g) * 0.1
    out = torch.zeros(M, 128, device="cuda", dtype=torch.bfloat16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = lambda t: C.c_void_p(t.data_ptr())
    rc = L.lpbox_mha_fused_dev(st, vp(X), vp(Wqkv), vp(Wo), vp(sc), vp(sh), vp(out), M, T)
    assert rc == 0, lpbox._capi.last_error()
    torch.cuda.synchronize()
    qkv = (X.float() @ Wqkv.float().t()).bfloat16().float().view(R, T, 3, 8, 16)
    q, k, v = qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2), qkv[:, :, 2].transpose(1, 2)      # (R, 8, T, 16)
    p = torch.softmax(q @ k.transpose(-1, -2) * 0.25, dim=-1).bfloat16().float()
    heads = (p @ v).transpose(1, 2).reshape(M, 128).bfloat16().float()
    ref = (X.float() + heads @ Wo.float().t()) * sc + sh
    err = (out.float() - ref).abs().max().item()
    assert err <= 0.03 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("kind,T", [("GraphAttentionEncoder", 20), ("MLPEncoder", 20), ("GraphAttentionEncoder", 5), ("GraphAttentionEncoder", 10)])
def test_policy_kernel_matches_torch_module(kind, T):
    from lpbox import policy
    from lpbox.policy_kernel import PolicyKernel
    torch.manual_seed(7)
    net = getattr(policy, kind)(tokens=T).cuda().eval()
    fill_deterministic(net)
    g = torch.Generator(device="cuda").manual_seed(T)
    x = torch.rand(3000, T, 5, device="cuda", generator=g)
    with torch.no_grad():
        ref = net(x)[1].reshape(-1)
    pk = PolicyKernel(net, chunk_rows=1024)
    got = pk(x)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 0.02
    assert pk.launch_count() > 0
