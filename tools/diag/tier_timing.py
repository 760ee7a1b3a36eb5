"""Diagnostic (not a test): step time and per-entry-point breakdown of the fp32 and bf16 tiers on config 2."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b200gat
from b200gat import synth, _lib
dev = torch.device("cuda:0")
nu, ni, n_inter, k = synth.CONFIGS["amazon"]
ei, feats = synth.make_graph(nu, ni, n_inter, k)
eid, fd = ei.to(dev), feats.to(dev)
u, i, j = (t.to(dev) for t in synth.make_triples(nu, ni, 200000))
tiers = os.environ.get("TIERS", "f32,bf16").split(",")
for name, dt in [t for t in (("f32", torch.float32), ("bf16", torch.bfloat16)) if t[0] in tiers]:
    torch.manual_seed(42)
    m = b200gat.PyGGAT(nu, ni, 128, 128, 2, heads=1, attn_dropout=0.1, feature_dtype=dt).to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
    def step():
        z = m(fd, eid); loss = b200gat.bpr_loss(z, nu, u, i, j); opt.zero_grad(set_to_none=True); loss.backward(); opt.step(); return loss
    for _ in range(4): step()
    torch.cuda.synchronize(); _lib.timing = {}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): l = step()
    b.record(); torch.cuda.synchronize()
    t, _lib.timing = _lib.timing, None
    print(name, "ms/step", round(a.elapsed_time(b) / 10, 3), "loss", l.item(), {k_.replace("b200gat_", ""): round(sum(x.elapsed_time(y) for x, y in v) / 10, 3) for k_, v in t.items()})
