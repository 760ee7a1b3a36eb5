"""Diagnostic: the three projection GEMM kernels alone at config-2 size (ncu target; prints CUDA-event times)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from b200gat import _lib as lib
dev = torch.device("cuda:0")
n = 690599
torch.manual_seed(0)
x = torch.randn(n, 128, device=dev); W = torch.randn(128, 128, device=dev) * 0.1
a_s, a_d = torch.randn(1, 128, device=dev), torch.randn(1, 128, device=dev)
h = torch.empty(n, 128, device=dev); s = torch.empty(n, 2, device=dev)
dh = torch.randn(n, 128, device=dev); ds = torch.randn(n, 2, device=dev)
dx = torch.empty(n, 128, device=dev); dW = torch.empty_like(W); da_s = torch.empty_like(a_s); da_d = torch.empty_like(a_d)
wsb = lib.dense_workspace_bytes(1, 128, 128); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
reps = int(os.environ.get("REPS", "5"))
def fwd():
    lib.call("b200gat_project_f32", lib.ptr(x), lib.ptr(W), lib.ptr(a_s), lib.ptr(a_d), n, 128, 1, 128, lib.ptr(h), lib.ptr(s), lib.ptr(ws), wsb, lib.stream())
def bwd():
    lib.call("b200gat_project_bwd_f32", lib.ptr(x), lib.ptr(W), lib.ptr(a_s), lib.ptr(a_d), lib.ptr(dh), lib.ptr(ds), n, 128, 1, 128,
             lib.ptr(dx), lib.ptr(dW), lib.ptr(da_s), lib.ptr(da_d), lib.ptr(ws), wsb, lib.stream())
for f in (fwd, bwd):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    print(f.__name__, round(a.elapsed_time(b) / reps, 4), "ms", flush=True)
