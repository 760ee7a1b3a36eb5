"""GPU: the tcgen05 (3xTF32) projection kernels against fp64 truth and against the CUDA-core fp32 kernels."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _project(lib, x, W, a_s, a_d, heads, mode):
    lib.set_gemm_mode(mode)
    n = x.shape[0]
    h = torch.empty(n, heads * 128, device=x.device)
    s = torch.empty(n, 2 * heads, device=x.device)
    wsb = lib.dense_workspace_bytes(heads, 128, 128)
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    lib.call("b200gat_project_f32", lib.ptr(x), lib.ptr(W), lib.ptr(a_s), lib.ptr(a_d), n, 128, heads, 128, lib.ptr(h), lib.ptr(s),
             lib.ptr(ws), wsb, lib.stream())
    torch.cuda.synchronize()
    return h, s


def _project_bwd(lib, x, W, a_s, a_d, dh, ds, mode):
    lib.set_gemm_mode(mode)
    n = x.shape[0]
    dh = dh.clone()
    dx = torch.empty(n, 128, device=x.device)
    dW, da_s, da_d = torch.empty_like(W), torch.empty_like(a_s), torch.empty_like(a_d)
    wsb = lib.dense_workspace_bytes(1, 128, 128)
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    lib.call("b200gat_project_bwd_f32", lib.ptr(x), lib.ptr(W), lib.ptr(a_s), lib.ptr(a_d), lib.ptr(dh), lib.ptr(ds), n, 128, 1, 128,
             lib.ptr(dx), lib.ptr(dW), lib.ptr(da_s), lib.ptr(da_d), lib.ptr(ws), wsb, lib.stream())
    torch.cuda.synchronize()
    return dx, dW, da_s, da_d


def _err(a, ref):
    return float((a.double().cpu() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("n", [1, 127, 128, 129, 5000, 148 * 128 * 3 + 77])
@pytest.mark.parametrize("heads", [1, 4])
def test_projection_forward(n, heads):
    from b200gat import _lib as lib
    dev = torch.device("cuda:0")
    torch.manual_seed(n + heads)
    x = torch.randn(n, 128, device=dev)
    W = torch.randn(heads * 128, 128, device=dev) * 0.1
    a_s, a_d = torch.randn(heads, 128, device=dev), torch.randn(heads, 128, device=dev)
    try:
        h_tc, s_tc = _project(lib, x, W, a_s, a_d, heads, lib.GEMM_TF32X3)
        h_32, s_32 = _project(lib, x, W, a_s, a_d, heads, lib.GEMM_FP32)
    finally:
        lib.set_gemm_mode(lib.GEMM_TF32X3)
    h_ref = x.double().cpu() @ W.double().cpu().t()
    hv = h_ref.view(n, heads, 128)
    s_ref = torch.cat([(hv * a_s.double().cpu()).sum(-1), (hv * a_d.double().cpu()).sum(-1)], dim=1)
    assert _err(h_32, h_ref) < 2e-6 and _err(s_32, s_ref) < 2e-6
    assert _err(h_tc, h_ref) < 3e-6, _err(h_tc, h_ref)        # 3xTF32: ~2^-21 relative
    assert _err(s_tc, s_ref) < 3e-6, _err(s_tc, s_ref)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 4097, 148 * 32 * 5 + 19, 300001])
def test_projection_backward(n):
    from b200gat import _lib as lib
    dev = torch.device("cuda:0")
    torch.manual_seed(n)
    x = torch.randn(n, 128, device=dev)
    W = torch.randn(128, 128, device=dev) * 0.1
    a_s, a_d = torch.randn(1, 128, device=dev), torch.randn(1, 128, device=dev)
    dh = torch.randn(n, 128, device=dev)
    ds = torch.randn(n, 2, device=dev)
    try:
        out_tc = _project_bwd(lib, x, W, a_s, a_d, dh, ds, lib.GEMM_TF32X3)
        out_32 = _project_bwd(lib, x, W, a_s, a_d, dh, ds, lib.GEMM_FP32)
    finally:
        lib.set_gemm_mode(lib.GEMM_TF32X3)
    x64, W64, dh64, ds64 = (t.double().cpu() for t in (x, W, dh, ds))
    dhf = dh64 + ds64[:, :1] * a_s.double().cpu() + ds64[:, 1:] * a_d.double().cpu()
    h64 = x64 @ W64.t()
    ref = (dhf @ W64, dhf.t() @ x64, (h64 * ds64[:, :1]).sum(0, keepdim=True), (h64 * ds64[:, 1:]).sum(0, keepdim=True))
    for name, a, b, r in zip(("dx", "dW", "da_src", "da_dst"), out_tc, out_32, ref):
        assert _err(b, r) < 1e-5, (name, "fp32", _err(b, r))
        assert _err(a, r) < 1e-5, (name, "tc", _err(a, r))


@pytest.mark.parametrize("n", [129, 20011])
@pytest.mark.parametrize("heads", [2, 4])
def test_projection_backward_multi_head(n, heads):
    """heads > 1: one tensor-core launch per head, dx accumulated over heads."""
    from b200gat import _lib as lib
    dev = torch.device("cuda:0")
    torch.manual_seed(n + heads)
    x = torch.randn(n, 128, device=dev)
    W = torch.randn(heads * 128, 128, device=dev) * 0.1
    a_s, a_d = torch.randn(heads, 128, device=dev), torch.randn(heads, 128, device=dev)
    dh = torch.randn(n, heads * 128, device=dev)
    ds = torch.randn(n, 2 * heads, device=dev)
    outs = {}
    try:
        for mode in (lib.GEMM_TF32X3, lib.GEMM_FP32):
            lib.set_gemm_mode(mode)
            dx = torch.empty(n, 128, device=dev)
            dW, da_s, da_d = torch.empty_like(W), torch.empty_like(a_s), torch.empty_like(a_d)
            wsb = lib.dense_workspace_bytes(heads, 128, 128)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            lib.call("b200gat_project_bwd_f32", lib.ptr(x), lib.ptr(W), lib.ptr(a_s), lib.ptr(a_d), lib.ptr(dh.clone()), lib.ptr(ds), n,
                     128, heads, 128, lib.ptr(dx), lib.ptr(dW), lib.ptr(da_s), lib.ptr(da_d), lib.ptr(ws), wsb, lib.stream())
            torch.cuda.synchronize()
            outs[mode] = (dx, dW, da_s, da_d)
    finally:
        lib.set_gemm_mode(lib.GEMM_TF32X3)
    x64, W64, dh64, ds64 = (t.double().cpu() for t in (x, W, dh, ds))
    a_s64, a_d64 = a_s.double().cpu(), a_d.double().cpu()
    dhf = dh64.view(n, heads, 128) + ds64[:, :heads, None] * a_s64 + ds64[:, heads:, None] * a_d64
    h64 = (x64 @ W64.t()).view(n, heads, 128)
    ref = (dhf.reshape(n, -1) @ W64, dhf.reshape(n, -1).t() @ x64, (h64 * ds64[:, :heads, None]).sum(0), (h64 * ds64[:, heads:, None]).sum(0))
    for name, a, b, r in zip(("dx", "dW", "da_src", "da_dst"), outs[lib.GEMM_TF32X3], outs[lib.GEMM_FP32], ref):
        assert _err(b, r) < 1e-5, (name, "fp32", _err(b, r))
        assert _err(a, r) < 1e-5, (name, "tc", _err(a, r))


def _bf16_round(t):
    return t.to(torch.bfloat16).double()


@pytest.mark.parametrize("n", [1, 129, 20011])
@pytest.mark.parametrize("f_in,channels,heads", [(128, 128, 1), (128, 128, 4), (256, 256, 4), (256, 128, 2), (128, 256, 1)])
def test_bf16_projection_forward_and_backward(n, f_in, channels, heads):
    """The bf16 tcgen05 GEMM (gemm_bf16.cu) at every supported shape -- K = 128 / 256 forward, K = heads * channels up to 1024 in
    the all-heads dx -- against fp64 arithmetic on the bf16-ROUNDED operands (that isolates the kernel from the rounding the
    tier allows), plus the fp32-accurate dW / da tiles."""
    from b200gat import _lib as lib
    dev = torch.device("cuda:0")
    torch.manual_seed(n + f_in + channels + heads)
    hc = heads * channels
    x = torch.randn(n, f_in, device=dev)
    W = torch.randn(hc, f_in, device=dev) * 0.1
    a_s, a_d = torch.randn(heads, channels, device=dev), torch.randn(heads, channels, device=dev)
    wsb = lib.dense_workspace_bytes(heads, channels, f_in)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    h = torch.empty(n, hc, dtype=torch.bfloat16, device=dev)
    s = torch.empty(n, 2 * heads, device=dev)
    lib.call("b200gat_project_bf16", lib.ptr(x), lib.ptr(W), lib.ptr(a_s), lib.ptr(a_d), n, f_in, heads, channels, lib.ptr(h), lib.ptr(s),
             lib.ptr(ws), wsb, lib.stream())
    torch.cuda.synchronize()
    xr, Wr = _bf16_round(x).cpu(), _bf16_round(W).cpu()
    h_ref = xr @ Wr.t()
    hv = h_ref.view(n, heads, channels)
    s_ref = torch.cat([(hv * a_s.double().cpu()).sum(-1), (hv * a_d.double().cpu()).sum(-1)], dim=1)
    assert _err(h.float(), h_ref) < 5e-3, _err(h.float(), h_ref)          # stored as bf16: 2^-9 relative
    assert _err(s, s_ref) < 1e-5, _err(s, s_ref)                          # taken from the fp32 accumulator
    dh = torch.randn(n, hc, device=dev)
    ds = torch.randn(n, 2 * heads, device=dev)
    dx = torch.empty(n, f_in, device=dev)
    dW, da_s, da_d = torch.empty_like(W), torch.empty_like(a_s), torch.empty_like(a_d)
    dh_before = dh.clone()
    lib.call("b200gat_project_bwd_bf16", lib.ptr(x), lib.ptr(W), lib.ptr(a_s), lib.ptr(a_d), lib.ptr(dh), lib.ptr(ds), n, f_in, heads,
             channels, lib.ptr(dx), lib.ptr(dW), lib.ptr(da_s), lib.ptr(da_d), lib.ptr(ws), wsb, lib.stream())
    torch.cuda.synchronize()
    assert torch.equal(dh, dh_before), "dh must not be modified"
    x64, W64, dh64, ds64 = (t.double().cpu() for t in (x, W, dh, ds))
    dhf = (dh64.view(n, heads, channels) + ds64[:, :heads, None] * a_s.double().cpu() + ds64[:, heads:, None] * a_d.double().cpu())
    dhf = dhf.reshape(n, hc)
    # the kernel rounds dh_full (formed in fp32) and W to bf16; an fp32-vs-fp64 difference in forming dh_full flips the bf16
    # rounding of a few elements (one bf16 ulp each), which shows up at ~1e-4 of max|dx| -- a wrong tile or K block would be O(1)
    dx_ref = _bf16_round(dhf.float()) @ Wr
    assert _err(dx, dx_ref) < 1e-3, _err(dx, dx_ref)
    assert _err(dx, dhf @ W64) < 2e-2                                       # and the bf16 tier's tolerance against unrounded truth
    h64 = (x64 @ W64.t()).view(n, heads, channels)
    assert _err(dW, dhf.t() @ x64) < 1e-5, _err(dW, dhf.t() @ x64)
    assert _err(da_s, (h64 * ds64[:, :heads, None]).sum(0)) < 1e-5
    assert _err(da_d, (h64 * ds64[:, heads:, None]).sum(0)) < 1e-5
