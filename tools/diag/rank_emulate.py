"""Diagnostic (not a test): one GPU stands in for rank 0 of a P-rank job (ShardedGAT(emulate=...), loopback fabric) and times
that rank's kernels of a config-2 training step.  Pull kernels run at HBM speed here, so only the compute kernels are
meaningful.   python tools/diag/rank_emulate.py [P ...] [--tier bf16] [--steps 10]"""
import os, sys, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
ap = argparse.ArgumentParser()
ap.add_argument("worlds", nargs="*", type=int, default=[1, 2, 4, 8])
ap.add_argument("--tier", default="f32")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--graph", action="store_true", help="replay the step from a CUDA graph as well")
args = ap.parse_args()
import b200gat
from b200gat import _lib, sharded, synth
dev = torch.device("cuda:0")
nu, ni, n_inter, k = synth.CONFIGS["amazon"]
ei, feats = synth.make_graph(nu, ni, n_inter, k)
u, i, j = (t.to(dev) for t in synth.make_triples(nu, ni, 200000))
for P in args.worlds:
    tr = sharded.ShardedGAT("pyg", nu, ni, feats, ei, hidden=128, layers=2, heads=1, attn_dropout=0.1, device=dev,
                            feature_dtype=torch.bfloat16 if args.tier == "bf16" else torch.float32, emulate=(0, P))
    for _ in range(3):
        tr.train_step(u, i, j)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tr.train_step(u, i, j)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    import time
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr.train_step(u, i, j)
    host_ms = (time.perf_counter() - t0) * 1e3 / args.steps      # host time to ENQUEUE a step (no sync inside)
    torch.cuda.synchronize()
    tr.fab.stats, _lib.timing = {}, {}
    for _ in range(args.steps):
        tr.train_step(u, i, j)
    torch.cuda.synchronize()
    stats, tr.fab.stats = tr.fab.stats, None
    timing, _lib.timing = _lib.timing, None
    kern = {k_.replace("b200gat_", ""): round(sum(a.elapsed_time(b) for a, b in v) / args.steps, 4) for k_, v in sorted(timing.items())}
    comm = {k_: round(sum(a.elapsed_time(b) for a, b in v) / args.steps, 4) for k_, v in stats.items()}
    print(json.dumps({"P": P, "tier": args.tier, "ms_step": round(ms, 3), "host_enqueue_ms": round(host_ms, 3), "n_loc": tr.n_loc,
                      "e_fwd": tr.g_fwd.n_edges, "kernels_ms": round(sum(kern.values()), 3), "loopback_comm_ms": round(sum(comm.values()), 3),
                      "kern": kern, "comm": comm}), flush=True)
    tr.close()
    del tr
    torch.cuda.empty_cache()
