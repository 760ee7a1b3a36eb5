"""Diagnostic (not a test): how fast is the forward edge kernel when the gathered rows fit in L2?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b200gat
from b200gat import synth, _lib
dev = torch.device("cuda:0")
nu, ni, n_inter, k = synth.CONFIGS["amazon"]
ei, _ = synth.make_graph(nu, ni, n_inter, k)
n = nu + ni
torch.manual_seed(0)
layer = b200gat.GATConv(128, 128, heads=1, concat=False, add_self_loops=False).to(dev).eval()
x = torch.randn(n, 128, device=dev)
def run(tag, e):
    eid = e.to(dev)
    _lib.timing = None
    with torch.no_grad():
        for _ in range(3): layer(x, eid)
        torch.cuda.synchronize(); _lib.timing = {}
        for _ in range(5): layer(x, eid)
        torch.cuda.synchronize()
    t, _lib.timing = _lib.timing, None
    ms = {k_: sum(a.elapsed_time(b) for a, b in v) / len(v) for k_, v in t.items()}
    print(tag, {k_.replace("b200gat_", ""): round(v, 3) for k_, v in ms.items()}, flush=True)
run("full graph (sources over 354 MB)", ei)
for frac in (4, 8):
    e2 = ei.clone(); e2[0] = e2[0] % (n // frac)          # same destinations/degrees, sources folded into 1/frac of the rows
    run(f"sources folded into {354 // frac} MB", e2)
# per-segment overhead: same edges, but every row cut into 4 segments (schedule with split=5 -> ~4 segments per row)
from b200gat.graph import build_graph
