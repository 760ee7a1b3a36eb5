"""Diagnostic: cosine kNN (k=20) time at the reference's item counts (63,001 interacted items; 498,196 catalogue)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import b200gat
from b200gat import _lib
dev = torch.device("cuda:0")
SIZES = ((63001, 128),) if os.environ.get("KNN_SIZES") == "one" else ((498196, 128),) if os.environ.get("KNN_SIZES") == "big" else ((63001, 128), (498196, 128)) if os.environ.get("KNN_SIZES") == "small" else ((63001, 128), (498196, 128), (63001, 384), (498196, 384))
for n, d in SIZES:
    g = torch.Generator(device="cpu").manual_seed(n)
    centers = torch.randn(n // 50, d, generator=g)
    emb = (centers[torch.randint(0, centers.shape[0], (n,), generator=g)] + 0.7 * torch.randn(n, d, generator=g)).to(dev)
    b200gat.knn_neighbors(emb, 20, 0.3); torch.cuda.synchronize(); stats = {}
    _lib.timing = {}
    t0 = time.perf_counter()
    idx, sim, counts = b200gat.knn_neighbors(emb, 20, 0.3, stats=stats)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t, _lib.timing = _lib.timing, None
    ms = sum(a.elapsed_time(b) for a, b in t["b200gat_knn_cosine_f32"])
    flops = 2.0 * n * n * d
    print(f"n={n} d={d}: {ms:.1f} ms device ({wall*1e3:.1f} ms wall), {flops/ms/1e9:.0f} TFLOP/s bf16-equivalent, edges kept {int(counts.sum())}, mean top sim {sim[:,0].mean().item():.3f}, exact-path rows {int(stats['exact_rows'].item())}", flush=True)
