"""Diagnostic (not a test): per-phase CUDA-event timing of the row-sharded step. Run under torchrun."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist

def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1: dist.init_process_group("nccl", device_id=dev)
    import b200gat
    from b200gat import _lib, sharded, synth
    nu, ni, n_inter, k = synth.CONFIGS["amazon"]
    ei, feats = synth.make_graph(nu, ni, n_inter, k)
    tr = sharded.ShardedGAT("pyg", nu, ni, feats, ei, hidden=128, layers=2, heads=1, attn_dropout=0.1, device=dev)
    u, i, j = (t.to(dev) for t in synth.make_triples(nu, ni, 200000))
    # wrap collectives with events
    coll = []
    orig = sharded.all_gather_rows
    def timed_gather(local_t, w):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = orig(local_t, w); b.record(); coll.append(("all_gather_rows", a, b)); return r
    sharded.all_gather_rows = timed_gather
    orig_pg = sharded.PeerExchange.gather
    def timed_pg(self_, b_, parts_):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = orig_pg(self_, b_, parts_); b.record(); coll.append(("peer_gather", a, b)); return r
    sharded.PeerExchange.gather = timed_pg
    orig_ar = dist.all_reduce
    def timed_ar(t, *a_, **k_):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = orig_ar(t, *a_, **k_); b.record(); coll.append(("all_reduce", a, b)); return r
    if world > 1: dist.all_reduce = timed_ar
    for _ in range(3): tr.train_step(u, i, j)
    torch.cuda.synchronize(); coll.clear(); _lib.timing = {}
    steps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1: dist.barrier()
    e0.record()
    for _ in range(steps): tr.train_step(u, i, j)
    e1.record(); torch.cuda.synchronize()
    res = {"rank": rank, "ms_step": e0.elapsed_time(e1) / steps, "n_loc": tr.n_loc, "e_fwd": tr.g_fwd.n_edges, "e_bwd": tr.g_bwd.n_edges}
    for name, evs in _lib.timing.items(): res[name] = round(sum(a.elapsed_time(b) for a, b in evs) / steps, 3)
    agg = {}
    for name, a, b in coll: agg[name] = agg.get(name, 0.0) + a.elapsed_time(b) / steps
    res.update({k_: round(v, 3) for k_, v in agg.items()})
    res["kernels+coll"] = round(sum(v for k_, v in res.items() if k_.startswith("b200gat_") or k_ in ("all_gather_rows", "all_reduce", "peer_gather")), 3)
    print(json.dumps(res), flush=True)
    if world > 1: dist.barrier(); dist.destroy_process_group()
main()
