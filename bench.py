#!/usr/bin/env python3
"""bench.py -- GAT fwd+bwd edges/sec on the BASELINE.json workloads (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5] [--loss bpr|bce]

``--config`` selects the BASELINE.json configuration (default 2, the one the metric is quoted on):
  1  custom GAT, 2 layers, d=128, BPR, 10k users / 20k items / 200k interactions + k=20 kNN (the reference's CPU-runnable case)
  2  PyG-dialect GATConv, 2 layers, d=128, heads=1, BPR, Amazon-Electronics shape (690,599 nodes, 13,342,152 edges)
  3  same graph, heads=4, bf16 projection, --loss bpr|bce
  4  full-graph item-embedding export: forward only (tools/export_item_embeddings.py:139-142), row-sharded at N GPUs
  5  power-law graph 10M users / 20M items / 200M interactions + k=20 kNN, 3 layers, d=256, heads=4, bf16 projection
     (needs >= 2 GPUs at full size; --scale S shrinks users/items/interactions by S for smaller boxes)

A training *step* is what the reference does on the device once per epoch (scripts/train_gat_pyg.py:305-322): one full-graph
forward, the ranking loss on S=200,000 sampled triples, one backward and one Adam step.  ``value`` = E * L / t_step
(edge-layer traversals per second, whole job), device-timed with CUDA events, inputs resident in HBM.  ``e2e`` = the same
step driven through the public module API from HOST buffers: the triples (config 4: item features and edge list) are copied
from pinned host memory and the loss (config 4: the item embeddings) is read back inside the timed region.

``--impl reference`` times the reference's CPU implementation of the same configuration on the host cores and never
imports the product package: the unmodified ``scripts.train_gat_custom.CustomGAT`` when /root/reference is present and the
configuration is the custom dialect (kind "reference"), otherwise the oracle's torch-CPU restatement (kind "port").
"""
from __future__ import annotations

import argparse
import datetime
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNIT = "edges/s"
S_TRIPLES = 200_000
FEAT_DIM = 128

# BASELINE.json configs[0..4]
CFGS = {
    1: dict(workload="cfg1", kind="custom", heads=1, hidden=128, layers=2, tier="f32", mode="train"),
    2: dict(workload="amazon", kind="pyg", heads=1, hidden=128, layers=2, tier="f32", mode="train"),
    3: dict(workload="amazon", kind="pyg", heads=4, hidden=128, layers=2, tier="bf16", mode="train"),
    4: dict(workload="amazon", kind="pyg", heads=1, hidden=128, layers=2, tier="f32", mode="export"),
    5: dict(workload="cfg5", kind="pyg", heads=4, hidden=256, layers=3, tier="bf16", mode="train"),
}


def load_synth():
    """The synthetic-workload generator, loaded from its file so that the CPU arms do not import the product package
    (importing ``b200gat`` maps libb200gat.so into the process)."""
    spec = importlib.util.spec_from_file_location("b200gat_synth_standalone",
                                                  os.path.join(ROOT, "plotpointe-gat-recommendation_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def resolve_cfg(args):
    cfg = dict(CFGS[args.config])
    cfg["id"] = args.config
    if args.workload:
        cfg["workload"] = args.workload
    if args.tier:
        cfg["tier"] = args.tier
    cfg["loss"] = args.loss
    cfg["scale"] = args.scale
    return cfg


def metric_name(cfg):
    return "gat_fwd_edges_per_sec" if cfg["mode"] == "export" else "gat_fwd_bwd_edges_per_sec"


def graph_dims(cfg, synth):
    nu, ni, n_inter, k = synth.CONFIGS[cfg["workload"]]
    if cfg["scale"] != 1:
        nu, ni, n_inter = (max(int(v / cfg["scale"]), 64) for v in (nu, ni, n_inter))
    return nu, ni, n_inter, k


def config_dict(cfg, nu, ni, n_inter, k, n_gpus):
    """The ``config`` object of the JSON line -- the same for both arms of a configuration."""
    e = 2 * n_inter + k * ni
    what = {"custom": "custom-dialect SimpleGATLayer", "pyg": "PyG-dialect GATConv"}[cfg["kind"]]
    if cfg["mode"] == "export":
        step = "forward only (item-embedding export, eval mode)"
    else:
        step = f"{cfg['loss'].upper()} on {S_TRIPLES} triples, train mode (attention dropout 0.1), Adam step"
    return {"workload": f"BASELINE config {cfg['id']} [{cfg['workload']}" + (f" / {cfg['scale']:g}" if cfg["scale"] != 1 else "") +
                        f"]: {what} x{cfg['layers']}, d={cfg['hidden']}, heads={cfg['heads']}, "
                        f"{'bf16 projection' if cfg['tier'] == 'bf16' else 'fp32'}, {step}; {nu} users, {ni} items, "
                        f"{n_inter} interactions + k={k} item kNN = {e} edges",
            "n_nodes": nu + ni, "n_edges": e, "layers": cfg["layers"], "heads": cfg["heads"], "hidden": cfg["hidden"],
            "l2": "per-step working set (h, x, dout: >= 3 x 354 MB) exceeds the 126 MB L2" if nu + ni > 300_000 else
                  "L2 flushed between timed steps (256 MB scratch write)",
            "parallelism": "1 GPU" if n_gpus == 1 else f"destination-row sharding over {n_gpus} GPUs"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md).

    A thread polls NVML (pynvml) every 2 ms from `start` on -- the calls release the GIL, so the launching thread is not held
    up -- and `begin` / `end` mark the timed region; `stop` reports the samples that fall inside it.  Without pynvml it falls
    back to one `nvidia-smi -lms 20` process whose timestamped lines are filtered the same way (coarser: nvidia-smi delivers
    a line every ~100 ms whatever the interval asked for)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.t_begin = self.t_end = None
        self.thread = None
        self.samples = []          # (perf_counter, sm_mhz, reasons bit mask)
        self.max_mhz = None
        self._stop = False

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.idx])
            except (ValueError, IndexError):
                return None
        return self.idx

    def _poll(self, pynvml, handle):
        while not self._stop:
            try:
                mhz = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                self.samples.append((time.perf_counter(), float(mhz), int(mask)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            phys = self._physical_index()
            handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(pynvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def begin(self):
        self.t_begin = datetime.datetime.now()
        self.p_begin = time.perf_counter()

    def end(self):
        self.t_end = datetime.datetime.now()
        self.p_end = time.perf_counter()

    def _summary(self, sel, window, source):
        sm, reasons = [r[0] for r in sel], set()
        for r in sel:
            reasons |= set(r[1])
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(sm), "window": window, "source": source}

    def stop(self):
        if self.thread is not None:
            p_stop = time.perf_counter()
            self._stop = True
            self.thread.join(timeout=1.0)
            t0 = getattr(self, "p_begin", p_stop)
            t1 = getattr(self, "p_end", p_stop)
            dec = lambda m: [n for n, bit in self.REASONS if m & bit]
            sel, window = [(mhz, dec(m)) for t, mhz, m in self.samples if t0 <= t <= t1], "timed region"
            if not sel:
                sel, window = [(mhz, dec(m)) for t, mhz, m in self.samples if t0 <= t <= p_stop], "timed region + end-to-end loop"
            return self._summary(sel, window, "NVML polled every 2 ms")
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither pynvml nor nvidia-smi available"]}
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
                names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                rows.append((ts, float(f[2]), float(f[3]), [n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        t0 = self.t_begin or t_stop
        t1 = self.t_end or t_stop
        if rows:
            self.max_mhz = max(r[2] for r in rows)
        sel, window = [(r[1], r[3]) for r in rows if t0 <= r[0] <= t1], "timed region"
        if not sel:
            sel, window = [(r[1], r[3]) for r in rows if t0 <= r[0] <= t_stop], "timed region + end-to-end loop (timed region shorter than one sample)"
        return self._summary(sel, window, "nvidia-smi -lms 20")


def algorithmic_bytes(n, e, h, c, dropout, row_bytes=4):
    """Gather-model bytes per launch of the two edge kernels.  row_bytes = bytes per
    element of the gathered matrix (4: fp32 tier, 2: bf16 tier)."""
    s = row_bytes
    fwd = e * (4 + 4 * h + h * c * s + (4 if dropout else 0)) + n * (16 + 4 * h + 4 * c + 8 * h)
    bwd = e * (4 + 16 * h + c * s + 4 * h + (4 if dropout else 0)) + n * (16 + h * c * s + h * c * 4 + 8 * h)
    sfx = "bf16" if row_bytes == 2 else "f32"
    return {f"b200gat_edge_fwd_{sfx}": fwd, f"b200gat_edge_bwd_{sfx}": bwd}


def roofline_from_timing(per_call, n, e, cfg, dropout):
    pk, pk_kind = peaks()
    bf16 = cfg["tier"] == "bf16"
    alg = algorithmic_bytes(n, e, cfg["heads"], cfg["hidden"], dropout=dropout, row_bytes=2 if bf16 else 4)
    edge_names = [k_ for k_ in alg if k_ in per_call]
    if not edge_names:
        return None
    dom = max(edge_names, key=lambda k_: per_call[k_]["ms_per_step"])
    ach = alg[dom] / (per_call[dom]["avg_ms"] * 1e-3) / 1e9
    roof = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": round(ach / pk["hbm_gbs"], 4), "traffic": None, "peak_source": pk_kind,
            "achieved_is": "gather-model (algorithmic) bytes / launch time: every gathered row is paid for, L2 hits included",
            "frac_of_nominal_8TBs": round(ach / 8000.0, 4),
            "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": round(per_call[dom]["avg_ms"], 4),
            "other_edge_kernels": {k_: {"achieved_gbs": round(alg[k_] / (per_call[k_]["avg_ms"] * 1e-3) / 1e9, 1),
                                        "frac": round(alg[k_] / (per_call[k_]["avg_ms"] * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
                                        "avg_launch_ms": round(per_call[k_]["avg_ms"], 4)} for k_ in edge_names if k_ != dom}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file) and cfg["id"] in (2, 4) and cfg["workload"] == "amazon" and cfg["tier"] == "f32":
        try:
            roof["traffic"] = json.load(open(traffic_file)).get(dom)
        except Exception:
            pass
    # `achieved` divides the gather-model bytes by the launch time; rows that hit in L2 make it exceed the DRAM peak.
    # `dram_frac` is the same launch time against the bytes DRAM really moved (ncu dram__bytes_read + write of the
    # committed --set full capture of this kernel on this workload).
    if roof["traffic"]:
        roof["dram_gbs"] = round(roof["traffic"] / (per_call[dom]["avg_ms"] * 1e-3) / 1e9, 1)
        roof["dram_frac"] = round(roof["dram_gbs"] / pk["hbm_gbs"], 4)
    if roof["frac"] > 1.0:
        roof["frac_note"] = ("frac > 1: the gather model charges every gathered row to HBM, but ~23 % of the gathered sectors hit in L2 "
                             "(ncu lts__t_sector_hit_rate); dram_frac is the same launch time against the bytes DRAM really moved")
    return roof


# ----------------------------------------------------------------------------------------------------- ours
def run_ours(args):
    cfg = resolve_cfg(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (ours) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import b200gat
    from b200gat import _lib, synth
    if world > 1 or cfg["id"] == 5:
        import torch.distributed as dist
        from b200gat import sharded
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        return sharded.bench_main(args, cfg, rank, world, dev)

    nu, ni, n_inter, k = graph_dims(cfg, synth)
    n = nu + ni
    ei, feats = synth.make_graph(nu, ni, n_inter, k)
    e = int(ei.shape[1])
    torch.manual_seed(42)
    bf16 = cfg["tier"] == "bf16"
    fdt = torch.bfloat16 if bf16 else torch.float32
    if cfg["kind"] == "custom":
        model = b200gat.CustomGAT(nu, ni, FEAT_DIM, cfg["hidden"], cfg["layers"], feature_dtype=fdt)
    else:
        model = b200gat.PyGGAT(nu, ni, FEAT_DIM, cfg["hidden"], cfg["layers"], heads=cfg["heads"], attn_dropout=0.1, feature_dtype=fdt)
    export = cfg["mode"] == "export"
    model = model.to(dev)
    model.eval() if export else model.train()
    opt = b200gat.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)   # same rule as the reference torch.optim.Adam
    loss_fn = b200gat.bpr_loss if cfg["loss"] == "bpr" else b200gat.bce_loss
    eid, fd = ei.to(dev), feats.to(dev)
    u, i, j = synth.make_triples(nu, ni, S_TRIPLES)
    hu, hi, hj = (t.pin_memory() for t in (u, i, j))
    du, di, dj = (t.to(dev) for t in (u, i, j))

    if export:
        def step(*_):
            with torch.no_grad():
                return model(fd, eid)
    else:
        def step(uu, ii, jj):
            z = model(fd, eid)
            loss = loss_fn(z, nu, uu, ii, jj)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss

    t0 = time.time()
    b200gat.graph_for(eid, n)
    torch.cuda.synchronize()
    graph_build_ms = (time.time() - t0) * 1e3
    small = n <= 300_000                     # the per-step working set would sit in L2: flush it between timed steps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None
    sampler = ClockSampler(local)
    sampler.start()                       # streams from here on; only the samples inside the timed region are kept
    for _ in range(args.warmup):
        step(du, di, dj)
    torch.cuda.synchronize()

    # ---- device-resident timed region ----------------------------------------------------------
    launches0 = _lib.launch_count()
    torch.cuda.synchronize()
    sampler.begin()
    if small:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evs:
            flush.zero_()
            a.record()
            step(du, di, dj)
            b.record()
        torch.cuda.synchronize()
        ms_step = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    else:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(args.steps):
            step(du, di, dj)
        ev[1].record()
        torch.cuda.synchronize()
        ms_step = ev[0].elapsed_time(ev[1]) / args.steps
    sampler.end()
    launches = _lib.launch_count() - launches0
    # ---- per-entry-point breakdown: a separate pass with CUDA events around every library call (not part of `value`)
    n_prof = min(args.steps, 10)
    _lib.timing = {}
    for _ in range(n_prof):
        if small:
            flush.zero_()
        step(du, di, dj)
    torch.cuda.synchronize()
    timing, _lib.timing = _lib.timing, None

    # ---- end to end: host buffers in, result out ------------------------------------------------------
    torch.cuda.synchronize()
    t_e2e = []
    if export:
        # the exporter's whole device job: features and edge list from the host, structure build, forward, item rows back
        hfeats, hei = feats.pin_memory(), ei.pin_memory()
        hout = torch.empty((ni, cfg["hidden"]), dtype=torch.float32).pin_memory()
        for _ in range(max(min(args.steps, 10), 3)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f2, e2 = hfeats.to(dev, non_blocking=True), hei.to(dev, non_blocking=True)
            with torch.no_grad():
                z = model(f2, e2)
            hout.copy_(z[nu:], non_blocking=True)
            torch.cuda.synchronize()
            t_e2e.append((time.perf_counter() - t0) * 1e3)
            del f2, e2, z
        h2d, d2h = int(feats.numel() * 4 + ei.numel() * 8), int(hout.numel() * 4)
        lv = float(hout[0, 0])
    else:
        for _ in range(max(args.steps, 3)):
            if small:
                flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            uu, ii, jj = (t.to(dev, non_blocking=True) for t in (hu, hi, hj))
            loss = step(uu, ii, jj)
            lv = loss.item()                         # device -> host read of the step's result
            t_e2e.append((time.perf_counter() - t0) * 1e3)
        h2d, d2h = int(3 * S_TRIPLES * 8), 4
    clocks = sampler.stop()
    e2e_ms = statistics.median(t_e2e)

    # ---- per-entry-point breakdown + roofline of the dominant kernel ------------------------------
    per_call = {}
    for name, evs_ in timing.items():
        ts = [a.elapsed_time(b) for a, b in evs_]
        per_call[name] = {"calls_per_step": len(ts) / n_prof, "avg_ms": sum(ts) / len(ts),
                          "ms_per_step": sum(ts) / n_prof}
    roof = roofline_from_timing(per_call, n, e, cfg, dropout=not export)
    extras = next_row_extras(model, fd, eid, nu, ni, dev) if (cfg["id"] == 2 and not args.no_next_rows) else None
    cpu = None
    if not args.no_cpu_baseline:
        del model, opt
        torch.cuda.empty_cache()
        cpu = cpu_arm(cfg, steps=1, warmup=1, budget_s=25.0)["cpu_baseline"]
    L = cfg["layers"]
    line = {
        "metric": metric_name(cfg), "value": e * L / (ms_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if bf16 else "f32", "data": "synthetic",
        "config": config_dict(cfg, nu, ni, n_inter, k, 1),
        "e2e": {"value": e * L / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
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
def _load_reference_module():
    """The reference's unmodified scripts/train_gat_custom.py (stand-in for its one missing import, google.cloud.storage),
    or None where the reference checkout does not exist (the GPU box)."""
    ref_root = os.environ.get("B200GAT_REFERENCE", "/root/reference")
    if not os.path.isfile(os.path.join(ref_root, "scripts", "train_gat_custom.py")):
        return None
    try:
        from oracle.make_golden import load_reference
        return load_reference()
    except Exception as exc:      # noqa: BLE001
        print(f"bench.py: reference import failed ({exc}); using the oracle port", file=sys.stderr)
        return None


def _edge_fraction(e, heads, hidden, factor=6.0):
    """Smallest power-of-two edge stride whose E x H x C fp32 intermediates (about `factor` of them live at the peak of a
    torch-CPU layer backward) fit in 60 % of the available host memory."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    frac = 1
    while factor * (e / frac) * heads * hidden * 4 > 0.6 * avail and frac < 4096:
        frac *= 2
    return frac, avail


def cpu_arm(cfg, steps, warmup, budget_s=None):
    """Times the reference's CPU implementation of configuration `cfg` on all host cores.  Returns the pieces of a JSON line.
    Never imports b200gat."""
    from oracle import gat_oracle as O
    synth = load_synth()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = dict(cfg)
    note = []
    if cfg["workload"] == "cfg5" and cfg["scale"] < 400:
        # 800 M edges x 1024 channels: no host runs that; the CPU arm runs the same generator 400 x smaller (nodes AND edges,
        # so the node-proportional work shrinks with the edges)
        cfg["scale"] = 400.0
        note.append("config 5 is infeasible on a host: the same generator at 1/400 scale (nodes and edges)")
    nu, ni, n_inter, k = graph_dims(cfg, synth)
    ei, feats = synth.make_graph(nu, ni, n_inter, k)
    u, i, j = synth.make_triples(nu, ni, S_TRIPLES)
    frac, avail = _edge_fraction(int(ei.shape[1]), cfg["heads"], cfg["hidden"])
    if frac > 1:
        ei = ei[:, ::frac].contiguous()
        note.append(f"every {frac}th edge over all nodes (the full edge set needs more than the {avail / 2**30:.0f} GiB of free host memory)")
    e = int(ei.shape[1])
    export = cfg["mode"] == "export"
    ref = _load_reference_module() if cfg["kind"] == "custom" else None
    torch.manual_seed(42)
    gen = torch.Generator().manual_seed(1234)
    loss_fn = O.bpr_loss if cfg["loss"] == "bpr" else O.bce_loss
    if ref is not None:
        kind = "reference"
        model = ref.CustomGAT(nu, ni, FEAT_DIM, cfg["hidden"], cfg["layers"])
        model.eval() if export else model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)       # train_gat_custom.py:335
        fwd = lambda: model(feats, ei)
        impl = "unmodified scripts/train_gat_custom.py CustomGAT (imported from the reference checkout) + its loss lines :350-359"
    else:
        kind = "port"
        init = O.init_custom_state if cfg["kind"] == "custom" else (lambda *a: O.init_pyg_state(*a, cfg["heads"]))
        params = {k_: v.clone().requires_grad_(not export) for k_, v in init(nu, ni, FEAT_DIM, cfg["hidden"], cfg["layers"]).items()}
        opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=1e-4)
        p = 0.0 if export else 0.1
        if cfg["kind"] == "custom":
            fwd = lambda: O.custom_gat_forward(params, feats, ei, p_drop=p, generator=gen)
        else:
            fwd = lambda: O.pyg_gat_forward(params, feats, ei, heads=cfg["heads"], p_drop=p, generator=gen)
        impl = "oracle/gat_oracle.py (torch-CPU restatement of the reference step)"

    def step():
        if export:
            with torch.no_grad():
                return float(fwd()[nu:].sum())
        z = fwd()
        loss = loss_fn(z, nu, u, i, j)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        lv = step()
        ts.append(time.perf_counter() - t0)
        if budget_s is not None and sum(ts) > budget_s:
            break
    t = statistics.median(ts)
    L = cfg["layers"]
    val = e * L / t
    what = "forward only" if export else f"fwd + {cfg['loss'].upper()} + bwd + Adam, train-mode dropout 0.1"
    sample = (f"{'the full graph' if frac == 1 and not note else '; '.join(note)}: {e} edges, {nu + ni} nodes, {S_TRIPLES} triples, {what}; "
              f"{impl}; {warmup} warm-up + {len(ts)} timed steps, median {t:.2f} s/step on {threads} threads")
    return {"value": val, "ms_per_step": t * 1e3, "loss": lv, "dims": (nu, ni, n_inter, k), "edges": e,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = resolve_cfg(args)
    synth = load_synth()
    res = cpu_arm(cfg, steps=args.steps, warmup=args.warmup)
    nu, ni, n_inter, k = graph_dims(cfg, synth)          # the configuration's own dimensions (config is shared with our arm)
    print(json.dumps({
        "impl": "reference", "metric": metric_name(cfg), "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg, nu, ni, n_inter, k, args.gpus),
        "cpu_baseline": res["cpu_baseline"],
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": res["loss"]}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CFGS), help="BASELINE.json configuration (default 2, the headline)")
    ap.add_argument("--loss", default="bpr", choices=["bpr", "bce"])
    ap.add_argument("--workload", default=None, help="override the configuration's graph (synth.CONFIGS name, e.g. tiny)")
    ap.add_argument("--tier", default=None, choices=["f32", "bf16"],
                    help="override the configuration's tier: f32 (rtol 1e-5) or bf16 projection (h / gathered dout stored as bf16, rtol 2e-2)")
    ap.add_argument("--scale", type=float, default=1.0, help="divide users / items / interactions of the workload by this")
    ap.add_argument("--stream-heads", default="auto", choices=["auto", "0", "1"],
                    help="heads > 1, bf16 tier, N-GPU path: one head at a time (config 5 at full size); auto = when [N, heads*C] per layer does not fit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
