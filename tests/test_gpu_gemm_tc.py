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


@pytest.mark.parametrize("n,n_peers", [(1, 1), (300, 3), (148 * 128 + 77, 7)])
def test_fused_projection_push_writes_identical_copies(n, n_peers):
    """b200gat_project_push_f32 / _bwd_push_f32 (the row-sharded path's fused projection + exchange): the GEMM epilogue stores
    every tile into the peers' buffers as well.  Here the "peers" are separate buffers of this device: the local result must
    be bit-identical to the plain entry point's, and every peer copy bit-identical to the local one (rows beyond n untouched)."""
    import ctypes
    from b200gat import _lib as lib
    dev = torch.device("cuda:0")
    torch.manual_seed(n)
    x = torch.randn(n, 128, device=dev)
    W = torch.randn(128, 128, device=dev) * 0.1
    a_s, a_d = torch.randn(1, 128, device=dev), torch.randn(1, 128, device=dev)
    h_ref, s_ref = _project(lib, x, W, a_s, a_d, 1, lib.GEMM_TF32X3)
    wsb = lib.dense_workspace_bytes(1, 128, 128)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    h, s = torch.empty_like(h_ref), torch.empty_like(s_ref)
    pad = 5                                                     # guard rows after the block: must stay untouched
    ph = [torch.full((n + pad, 128), -7.0, device=dev) for _ in range(n_peers)]
    ps = [torch.full((n + pad, 2), -7.0, device=dev) for _ in range(n_peers)]
    arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    lib.call("b200gat_project_push_f32", lib.ptr(x), lib.ptr(W), lib.ptr(a_s), lib.ptr(a_d), n, 128, 1, 128, lib.ptr(h), lib.ptr(s),
             arr(ph), arr(ps), n_peers, lib.ptr(ws), wsb, lib.stream())
    torch.cuda.synchronize()
    assert torch.equal(h, h_ref) and torch.equal(s, s_ref)
    for q in range(n_peers):
        assert torch.equal(ph[q][:n], h) and torch.equal(ps[q][:n], s)
        assert bool((ph[q][n:] == -7.0).all()) and bool((ps[q][n:] == -7.0).all())
    # backward: dx into the peers' buffers
    dh = torch.randn(n, 128, device=dev)
    ds = torch.randn(n, 2, device=dev)
    dx_ref, dW_ref, das_ref, dad_ref = _project_bwd(lib, x, W, a_s, a_d, dh, ds, lib.GEMM_TF32X3)
    dx = torch.empty(n, 128, device=dev)
    dW, da_s2, da_d2 = torch.empty_like(W), torch.empty_like(a_s), torch.empty_like(a_d)
    pd = [torch.full((n + pad, 128), -7.0, device=dev) for _ in range(n_peers)]
    lib.call("b200gat_project_bwd_push_f32", lib.ptr(x), lib.ptr(W), lib.ptr(a_s), lib.ptr(a_d), lib.ptr(dh), lib.ptr(ds), n, 128, 1, 128,
             lib.ptr(dx), arr(pd), n_peers, lib.ptr(dW), lib.ptr(da_s2), lib.ptr(da_d2), lib.ptr(ws), wsb, lib.stream())
    torch.cuda.synchronize()
    assert torch.equal(dx, dx_ref) and torch.equal(dW, dW_ref) and torch.equal(da_s2, das_ref) and torch.equal(da_d2, dad_ref)
    for q in range(n_peers):
        assert torch.equal(pd[q][:n], dx) and bool((pd[q][n:] == -7.0).all())


def test_fused_projection_push_rejects_other_shapes():
    from b200gat import _lib as lib
    dev = torch.device("cuda:0")
    x = torch.randn(8, 128, device=dev)
    W = torch.randn(256, 128, device=dev)
    a = torch.randn(2, 128, device=dev)
    h, s = torch.empty(8, 256, device=dev), torch.empty(8, 4, device=dev)
    wsb = lib.dense_workspace_bytes(2, 128, 128)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    with pytest.raises(RuntimeError, match="fused projection"):
        lib.call("b200gat_project_push_f32", lib.ptr(x), lib.ptr(W), lib.ptr(a), lib.ptr(a), 8, 128, 2, 128, lib.ptr(h), lib.ptr(s),
                 None, None, 0, lib.ptr(ws), wsb, lib.stream())


@pytest.mark.parametrize("world", [2, 3, 8])
def test_peer_push_and_pull_gathers_in_one_process(world):
    """The exchange kernels with every "rank" played by this process on one device (separate buffers, launches in stream
    order, so no kernel ever waits for a later one): after every rank has pushed (or pulled), all buffers hold all blocks;
    the reduce-pull sums the ranks' values in rank order."""
    import ctypes
    from b200gat import _lib as lib
    dev = torch.device("cuda:0")
    fb = ctypes.c_size_t(0)
    lib._check(lib._lib.b200gat_peer_flag_bytes(4, ctypes.byref(fb)), "flag_bytes")
    blk0, blk1 = 50_000 * 16, 1_000 * 16                     # two parts (bytes per rank block), 16-byte multiples
    off0 = fb.value
    off1 = off0 + world * blk0
    off2 = off1 + world * blk1                                # a reduce region of 1000 floats
    total = off2 + 4000
    torch.manual_seed(world)
    for mode in ("push", "pull"):
        bufs = [torch.zeros(total, dtype=torch.uint8, device=dev) for _ in range(world)]
        truth0 = [torch.randint(0, 255, (blk0,), dtype=torch.uint8, device=dev) for _ in range(world)]
        truth1 = [torch.randint(0, 255, (blk1,), dtype=torch.uint8, device=dev) for _ in range(world)]
        red = [torch.randn(1000, device=dev) for _ in range(world)]
        for r in range(world):
            bufs[r][off0 + r * blk0:off0 + (r + 1) * blk0] = truth0[r]
            bufs[r][off1 + r * blk1:off1 + (r + 1) * blk1] = truth1[r]
            bufs[r][off2:off2 + 4000] = red[r].view(torch.uint8)
        bases = (ctypes.c_void_p * world)(*[b.data_ptr() for b in bufs])
        offs, blks = (ctypes.c_uint64 * 2)(off0, off1), (ctypes.c_uint64 * 2)(blk0, blk1)
        st = lib.stream()
        if mode == "push":
            for r in range(world):
                lib._check(lib._lib.b200gat_peer_push(bases, world, r, 2, offs, blks, st), "push")
        else:
            for r in range(world):                              # every rank signals first: the pulls then never spin
                lib._check(lib._lib.b200gat_peer_signal(bases, world, r, 0, 1, st), "signal")
            for r in range(world):
                lib._check(lib._lib.b200gat_peer_allgather(bases, world, r, 0, 1, 2, offs, blks, st), "allgather")
        outs = [torch.empty(1000, device=dev) for _ in range(world)]
        for r in range(world):
            lib._check(lib._lib.b200gat_peer_signal(bases, world, r, 1, 1, st), "signal")
        for r in range(world):
            lib._check(lib._lib.b200gat_peer_reduce_f32(bases, world, r, 1, 1, off2, 0, 1000, lib.ptr(outs[r]), st), "reduce")
        torch.cuda.synchronize()
        expect = torch.zeros(1000, device=dev)
        for r in range(world):
            expect = expect + red[r]                            # rank order, fp32
        for r in range(world):
            assert torch.equal(bufs[r][off0:off1], torch.cat(truth0)), (mode, r)
            assert torch.equal(bufs[r][off1:off2], torch.cat(truth1)), (mode, r)
            assert torch.equal(outs[r], expect), (mode, r)
