#!/usr/bin/env python3
"""bench.py -- GAT fwd+bwd edges/sec on the BASELINE.json workload (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload amazon|cfg1|tiny]

A *step* is what the reference does on the device once per epoch (scripts/train_gat_pyg.py:305-322): one full-graph
forward of the 2-layer GAT, the BPR loss on S=200,000 sampled triples, one backward and one Adam step.
``value`` = E * L / t_step (edge-layer traversals per second, whole job), device-timed with CUDA events, inputs
resident in HBM.  ``e2e`` = the same step driven through the public module API from HOST buffers: the triples are
copied from pinned host memory and the loss is read back inside the timed region.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "gat_fwd_bwd_edges_per_sec"
UNIT = "edges/s"
S_TRIPLES = 200_000
LAYERS = 2
HIDDEN = 128
HEADS = 1


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md).

    nvidia-smi needs ~0.2 s before its first line, longer than a short timed region, so it is started before the warm-up
    (`start`) and keeps streaming one timestamped line every 20 ms; `begin` / `end` mark the timed region on the host
    clock and `stop` keeps the samples whose timestamps fall inside it.  If the region was too short to catch one, the
    samples of the end-to-end loop that follows (same load) are used and `window` says so."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def begin(self):
        self.t_begin = datetime.datetime.now()

    def end(self):
        self.t_end = datetime.datetime.now()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_stop = datetime.datetime.now()
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(f[2]), float(f[3]), f[5:9]))
            except ValueError:
                continue
        t0 = self.t_begin or t_stop
        t1 = self.t_end or t_stop
        sel, window = [r for r in rows if t0 <= r[0] <= t1], "timed region"
        if not sel:
            sel, window = [r for r in rows if t0 <= r[0] <= t_stop], "timed region + end-to-end loop (timed region shorter than one sample)"
        sm, mx, reasons = [r[1] for r in sel], [r[2] for r in sel], set()
        for r in sel:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def algorithmic_bytes(n, e, h, c, dropout, row_bytes=4):
    """Gather-model bytes per launch of the two edge kernels.  row_bytes = bytes per
    element of the gathered matrix (4: fp32 tier, 2: bf16 tier)."""
    s = row_bytes
    fwd = e * (4 + 4 * h + h * c * s + (4 if dropout else 0)) + n * (16 + 4 * h + 4 * c + 8 * h)
    bwd = e * (4 + 16 * h + c * s + 4 * h + (4 if dropout else 0)) + n * (16 + h * c * s + h * c * 4 + 8 * h)
    sfx = "bf16" if row_bytes == 2 else "f32"
    return {f"b200gat_edge_fwd_{sfx}": fwd, f"b200gat_edge_bwd_{sfx}": bwd}


# ----------------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import b200gat
    from b200gat import _lib, synth
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (ours) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from b200gat import sharded
        return sharded.bench_main(args, rank, world, dev)

    nu, ni, n_inter, k = synth.CONFIGS[args.workload]
    n = nu + ni
    ei, feats = synth.make_graph(nu, ni, n_inter, k)
    e = int(ei.shape[1])
    torch.manual_seed(42)
    bf16 = args.tier == "bf16"
    model = b200gat.PyGGAT(nu, ni, 128, HIDDEN, LAYERS, heads=HEADS, attn_dropout=0.1,
                           feature_dtype=torch.bfloat16 if bf16 else torch.float32).to(dev).train()
    opt = b200gat.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)   # same rule as the reference torch.optim.Adam
    eid, fd = ei.to(dev), feats.to(dev)
    u, i, j = synth.make_triples(nu, ni, S_TRIPLES)
    hu, hi, hj = (t.pin_memory() for t in (u, i, j))
    du, di, dj = (t.to(dev) for t in (u, i, j))

    def step(uu, ii, jj):
        z = model(fd, eid)
        loss = b200gat.bpr_loss(z, nu, uu, ii, jj)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    t0 = time.time()
    b200gat.graph_for(eid, n)
    torch.cuda.synchronize()
    graph_build_ms = (time.time() - t0) * 1e3
    sampler = ClockSampler(local)
    sampler.start()                       # streams from here on; only the samples inside the timed region are kept
    for _ in range(args.warmup):
        step(du, di, dj)
    torch.cuda.synchronize()

    # ---- device-resident timed region ----------------------------------------------------------
    launches0 = _lib.launch_count()
    _lib.timing = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    sampler.begin()
    ev[0].record()
    for _ in range(args.steps):
        loss = step(du, di, dj)
    ev[1].record()
    torch.cuda.synchronize()
    sampler.end()
    timing, _lib.timing = _lib.timing, None
    launches = _lib.launch_count() - launches0
    ms_step = ev[0].elapsed_time(ev[1]) / args.steps

    # ---- end to end: host triples in, loss out ----------------------------------------------------
    torch.cuda.synchronize()
    t_e2e = []
    for _ in range(max(args.steps, 3)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        uu, ii, jj = (t.to(dev, non_blocking=True) for t in (hu, hi, hj))
        loss = step(uu, ii, jj)
        lv = loss.item()                         # device -> host read of the step's result
        t_e2e.append((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    e2e_ms = statistics.median(t_e2e)

    # ---- per-entry-point breakdown + roofline of the dominant kernel ------------------------------
    per_call = {}
    for name, evs in timing.items():
        ts = [a.elapsed_time(b) for a, b in evs]
        per_call[name] = {"calls_per_step": len(ts) / args.steps, "avg_ms": sum(ts) / len(ts),
                          "ms_per_step": sum(ts) / args.steps}
    pk, pk_kind = peaks()
    alg = algorithmic_bytes(n, e, HEADS, HIDDEN, dropout=True, row_bytes=2 if bf16 else 4)
    edge_names = [k_ for k_ in alg if k_ in per_call]
    dom = max(edge_names, key=lambda k_: per_call[k_]["ms_per_step"])
    ach = alg[dom] / (per_call[dom]["avg_ms"] * 1e-3) / 1e9
    roof = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": round(ach / pk["hbm_gbs"], 4), "traffic": None, "peak_source": pk_kind,
            "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": round(per_call[dom]["avg_ms"], 4),
            "other_edge_kernels": {k_: {"achieved_gbs": round(alg[k_] / (per_call[k_]["avg_ms"] * 1e-3) / 1e9, 1),
                                        "frac": round(alg[k_] / (per_call[k_]["avg_ms"] * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
                                        "avg_launch_ms": round(per_call[k_]["avg_ms"], 4)} for k_ in edge_names if k_ != dom}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roof["traffic"] = json.load(open(traffic_file)).get(dom)
        except Exception:
            pass
    # `achieved` divides the gather-model (algorithmic) bytes, which pay for every gathered row, by the launch time; rows
    # that hit in L2 make it exceed the DRAM peak.  `dram_frac` is the same launch time against the bytes DRAM really
    # moved (ncu dram__bytes_read + write of the committed --set full capture).
    if roof["traffic"]:
        roof["dram_gbs"] = round(roof["traffic"] / (per_call[dom]["avg_ms"] * 1e-3) / 1e9, 1)
        roof["dram_frac"] = round(roof["dram_gbs"] / pk["hbm_gbs"], 4)

    extras = next_row_extras(model, fd, eid, nu, ni, dev)
    cpu = cpu_baseline(args, nu, ni, ei, feats, (u, i, j)) if not args.no_cpu_baseline else None
    line = {
        "metric": METRIC, "value": e * LAYERS / (ms_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if bf16 else "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: PyG-dialect GATConv x{LAYERS}, d={HIDDEN}, heads={HEADS}, BPR on "
                               f"{S_TRIPLES} triples, train mode (attention dropout 0.1), Adam step; {nu} users, {ni} items, "
                               f"{n_inter} interactions + k={k} item kNN = {e} edges",
                   "n_nodes": n, "n_edges": e, "layers": LAYERS, "l2": "per-step working set (h, x, dout: 3 x 354 MB) exceeds the 126 MB L2",
                   "parallelism": "1 GPU"},
        "e2e": {"value": e * LAYERS / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(3 * S_TRIPLES * 8), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "epoch_time_ms": ms_step, "graph_build_ms": round(graph_build_ms, 2), "loss": lv,
        "breakdown_ms_per_step": {k_: round(v["ms_per_step"], 4) for k_, v in sorted(per_call.items())},
        "next_rows": extras,
    }
    print(json.dumps(line))


def next_row_extras(model, fd, eid, nu, ni, dev, n_eval=20000, neg_k=1000):
    """Measurement of the widened rows: sampled-evaluation ranks, 1 positive + 1000 negatives per user
    as in the reference's eval_sampled (eval_neg_k = 1000), device time with CUDA events; the oracle on a small sample."""
    import b200gat
    from oracle import gat_oracle as O
    g = torch.Generator().manual_seed(7)
    users = torch.randint(0, nu, (n_eval,), generator=g)
    cands = torch.randint(0, ni, (n_eval, neg_k + 1), generator=g)
    model.eval()
    with torch.no_grad():
        z = model(fd, eid)
    model.train()
    ud, cd = users.to(dev), cands.to(dev)
    b200gat.eval_ranks(z, nu, ud, cd)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        ranks = b200gat.eval_ranks(z, nu, ud, cd)
    b_.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b_) / 3
    zc = z.cpu()
    ns = 500
    t0 = time.perf_counter()
    ref_ranks, _ = O.eval_ranks(zc, nu, users[:ns], cands[:ns])
    cpu_s = time.perf_counter() - t0
    agree = float((ranks[:ns].cpu().long() == ref_ranks).float().mean())
    gather_bytes = n_eval * (neg_k + 2) * z.shape[1] * 4
    # f1: cosine kNN (k=20, min_similarity 0.3) at the reference's own item count (63,001 interacted items,
    # PHASE0_REPORT.md:190-193: 77.91 s on an n1-highmem-8); oracle restatement on a 6,000-item sample
    gk = torch.Generator().manual_seed(11)
    nk = 63001
    centers = torch.randn(nk // 50, 128, generator=gk)
    emb = centers[torch.randint(0, centers.shape[0], (nk,), generator=gk)] + 0.7 * torch.randn(nk, 128, generator=gk)
    embd = emb.to(dev)
    b200gat.knn_neighbors(embd, 20, 0.3)
    torch.cuda.synchronize()
    a.record()
    for _ in range(3):
        _, _, counts = b200gat.knn_neighbors(embd, 20, 0.3)
    b_.record()
    torch.cuda.synchronize()
    knn_ms = a.elapsed_time(b_) / 3
    t0 = time.perf_counter()
    O.build_ii_knn(emb[:6000].numpy(), 20, 0.3)
    knn_cpu_s = time.perf_counter() - t0
    knn = {"items": nk, "k": 20, "ms": round(knn_ms, 2), "item_pairs_per_s": nk * nk / (knn_ms * 1e-3),
           "tflops_bf16_equivalent": round(2.0 * nk * nk * 128 / (knn_ms * 1e-3) / 1e12, 1), "edges": int(counts.sum()),
           "cpu_oracle_item_pairs_per_s": round(6000 * 6000 / knn_cpu_s, 1), "cpu_sample": "first 6000 items, numpy restatement",
           "reference_published_s": 77.91}
    return {"knn": knn, "eval_ranks": {"users_per_s": n_eval / (ms * 1e-3), "ms": round(ms, 3), "n_users": n_eval, "candidates": neg_k + 1,
                           "achieved_gbs": round(gather_bytes / (ms * 1e-3) / 1e9, 1),
                           "cpu_oracle_users_per_s": round(ns / cpu_s, 1), "rank_agreement_with_oracle": agree}}


# ------------------------------------------------------------------------------------------------ CPU arms
def _oracle_step_fn(nu, ni, ei, feats, triples, threads):
    """The oracle's restatement of the reference step (PyG dialect, eval-mode dropout) on the host cores."""
    from oracle import gat_oracle as O
    import b200gat
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    m = b200gat.PyGGAT(nu, ni, 128, HIDDEN, LAYERS, heads=HEADS, attn_dropout=0.1)   # parameters only (CPU tensors)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=1e-4)
    u, i, j = triples

    def step():
        z = O.pyg_gat_forward(params, feats, ei, heads=HEADS)
        loss = O.bpr_loss(z, nu, u, i, j)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def _sample_graph(ei, frac):
    """Bounded sample of the workload: every `frac`-th edge of the same graph over the full node set."""
    return ei[:, ::frac].contiguous()


def cpu_baseline(args, nu, ni, ei, feats, triples, steps=2, warmup=1):
    threads = os.cpu_count() or 1
    frac = 8 if args.workload == "amazon" else 1
    eis = _sample_graph(ei, frac)
    step = _oracle_step_fn(nu, ni, eis, feats, triples, threads)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return {"value": int(eis.shape[1]) * LAYERS / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"every {frac}th edge of the same graph ({int(eis.shape[1])} edges, all {nu + ni} nodes), same "
                      f"{S_TRIPLES} triples, fwd+BPR+bwd+Adam, oracle/gat_oracle.py (torch CPU), {warmup} warm-up + "
                      f"{steps} timed steps, median {t:.2f} s/step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from b200gat import synth
    nu, ni, n_inter, k = synth.CONFIGS[args.workload]
    ei, feats = synth.make_graph(nu, ni, n_inter, k)
    triples = synth.make_triples(nu, ni, S_TRIPLES)
    threads = os.cpu_count() or 1
    frac = 8 if args.workload == "amazon" else 1
    eis = _sample_graph(ei, frac)
    step = _oracle_step_fn(nu, ni, eis, feats, triples, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lv = step()
    ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e = int(eis.shape[1])
    val = e * LAYERS / (ms * 1e-3)
    sample = (f"every {frac}th edge of the {args.workload} graph ({e} edges, all {nu + ni} nodes), {S_TRIPLES} triples, "
              f"fwd+BPR+bwd+Adam per step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} (CPU arm: {sample})", "n_edges": e, "layers": LAYERS},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": lv}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="amazon", choices=["amazon", "cfg1", "tiny"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tier", default="f32", choices=["f32", "bf16"],
                    help="f32 (headline, rtol 1e-5 tier) or bf16 projection (h / gathered dout stored as bf16, rtol 2e-2 tier)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
