"""Diagnostic: BASELINE config 3 (heads=4, d=128, bf16 projection, BCE vs BPR) step time and breakdown."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b200gat
from b200gat import synth, _lib
dev = torch.device("cuda:0")
nu, ni, n_inter, k = synth.CONFIGS["amazon"]
ei, feats = synth.make_graph(nu, ni, n_inter, k)
eid, fd = ei.to(dev), feats.to(dev)
u, i, j = (t.to(dev) for t in synth.make_triples(nu, ni, 200000))
cases = (("h4 f32 bpr", torch.float32, b200gat.bpr_loss), ("h4 bf16 bpr", torch.bfloat16, b200gat.bpr_loss), ("h4 bf16 bce", torch.bfloat16, b200gat.bce_loss))
only = os.environ.get("CASES")
for name, dt, loss_fn in [c for c in cases if not only or c[0] in only.split(",")]:
    torch.manual_seed(42)
    m = b200gat.PyGGAT(nu, ni, 128, 128, 2, heads=4, attn_dropout=0.1, feature_dtype=dt).to(dev).train()
    opt = b200gat.Adam(m.parameters(), lr=1e-3, weight_decay=1e-4)
    def step():
        z = m(fd, eid); loss = loss_fn(z, nu, u, i, j); opt.zero_grad(set_to_none=True); loss.backward(); opt.step(); return loss
    for _ in range(3): step()
    torch.cuda.synchronize(); _lib.timing = {}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): l = step()
    b.record(); torch.cuda.synchronize()
    t, _lib.timing = _lib.timing, None
    print(name, "ms/step", round(a.elapsed_time(b) / 5, 3), "loss", round(l.item(), 6), "peak GB", round(torch.cuda.max_memory_allocated() / 1e9, 1),
          {k_.replace("b200gat_", ""): round(sum(x.elapsed_time(y) for x, y in v) / 5, 3) for k_, v in t.items()}, flush=True)
    del m, opt
