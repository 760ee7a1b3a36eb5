#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (test infrastructure).

Runs only in the build container, where /root/reference exists.  The reference's
``scripts/train_gat_custom.py`` imports ``google.cloud.storage`` at module scope (:20); that package
is not installed, so an empty stand-in module is placed in ``sys.modules`` first.  Nothing of the
reference is copied: its classes are imported, executed on seeded inputs, and only the resulting
tensors are stored.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz
    python oracle/make_golden.py knn        # only the kNN fixtures (runs graphs/build_ii_knn.py as a script)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("B200GAT_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    for name in ("google", "google.cloud", "google.cloud.storage"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["google.cloud.storage"].Client = object
    sys.modules["google"].cloud = sys.modules["google.cloud"]
    sys.modules["google.cloud"].storage = sys.modules["google.cloud.storage"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import scripts.train_gat_custom as ref  # noqa: WPS433
    return ref


def small_graph(rng, n_nodes, n_edges, n_isolated=5, hub=None):
    """Random multigraph with duplicate edges, some nodes without in-edges, optional hub row."""
    src = rng.integers(0, n_nodes, size=n_edges)
    dst = rng.integers(0, n_nodes - n_isolated, size=n_edges)   # last n_isolated nodes: no in-edges
    if hub is not None:
        dst[: n_edges // 4] = hub
    src[10:20] = src[0:10]
    dst[10:20] = dst[0:10]                                       # duplicated edges
    return torch.from_numpy(np.stack([src, dst]).astype(np.int64))


def layer_case(ref, seed, n_nodes, n_edges, dim, x_scale, hub=None):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    layer = ref.SimpleGATLayer(dim, dim).double().eval()
    with torch.no_grad():           # widen the attention vectors so some logits leave [-10, 10]
        layer.a_src.mul_(x_scale)
        layer.a_dst.mul_(x_scale)
    ei = small_graph(rng, n_nodes, n_edges, hub=hub)
    x64 = torch.randn(n_nodes, dim, dtype=torch.float64)
    g64 = torch.randn(n_nodes, dim, dtype=torch.float64)
    out = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        lay = ref.SimpleGATLayer(dim, dim).to(dt).eval()
        lay.load_state_dict({k: v.to(dt) for k, v in layer.state_dict().items()})
        x = x64.to(dt).clone().requires_grad_(True)
        y = lay(x, ei)
        (y * g64.to(dt)).sum().backward()
        out[f"out_{tag}"] = y.detach().numpy()
        out[f"dx_{tag}"] = x.grad.numpy()
        out[f"dW_{tag}"] = lay.lin.weight.grad.numpy()
        out[f"da_src_{tag}"] = lay.a_src.grad.numpy()
        out[f"da_dst_{tag}"] = lay.a_dst.grad.numpy()
    out.update(edge_index=ei.numpy(), x=x64.numpy(), g=g64.numpy(), W=layer.lin.weight.detach().numpy(),
               a_src=layer.a_src.detach().numpy(), a_dst=layer.a_dst.detach().numpy())
    return out


def model_case(ref, seed, n_users, n_items, feat_dim, hidden, layers, n_inter, n_triples):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    train_pos = {}
    for u in range(n_users):
        k = int(rng.integers(1, 2 * n_inter // n_users))
        train_pos[u] = rng.integers(0, n_items, size=k)          # duplicates allowed
    ei = ref.build_edge_index(n_users, n_items, train_pos)
    # item-item block appended after the U-I block
    ii_src = n_users + rng.integers(0, n_items, size=4 * n_items)
    ii_dst = n_users + rng.integers(0, n_items, size=4 * n_items)
    ei = torch.cat([ei, torch.from_numpy(np.stack([ii_src, ii_dst]).astype(np.int64))], dim=1)
    feats = torch.nn.functional.normalize(torch.randn(n_items, feat_dim, dtype=torch.float64), dim=1)
    u = torch.from_numpy(rng.integers(0, n_users, size=n_triples))
    i = torch.from_numpy(rng.integers(0, n_items, size=n_triples))
    j = torch.from_numpy(rng.integers(0, n_items, size=n_triples))
    base = ref.CustomGAT(n_users, n_items, feat_dim, hidden, layers).double().eval()
    out = {"edge_index": ei.numpy(), "item_feats": feats.numpy(), "u": u.numpy(), "i": i.numpy(), "j": j.numpy(),
           "n_users": np.int64(n_users), "n_items": np.int64(n_items),
           "ui_users": np.concatenate([np.full(len(v), k) for k, v in train_pos.items()]),
           "ui_items": np.concatenate(list(train_pos.values()))}
    for k, v in base.state_dict().items():
        out["param:" + k] = v.numpy()
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        for loss_name in ("bpr", "bce"):
            m = ref.CustomGAT(n_users, n_items, feat_dim, hidden, layers).to(dt).eval()
            m.load_state_dict({k: v.to(dt) for k, v in base.state_dict().items()})
            z = m(feats.to(dt), ei)
            uu, ii = z[:n_users], z[n_users:]
            pos = (uu[u] * ii[i]).sum(dim=-1)               # the reference's loss lines, executed here
            neg = (uu[u] * ii[j]).sum(dim=-1)               # through the reference's own model output
            if loss_name == "bpr":
                loss = -torch.log(torch.sigmoid(pos - neg) + 1e-8).mean()
            else:
                logits = torch.cat([pos, neg], dim=0)
                labels = torch.cat([torch.ones_like(pos), torch.zeros_like(neg)], dim=0)
                loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, labels)
            loss.backward()
            out[f"z_{tag}"] = z.detach().numpy()
            out[f"loss_{loss_name}_{tag}"] = loss.detach().numpy()
            for k, p in m.named_parameters():
                out[f"grad_{loss_name}_{tag}:{k}"] = p.grad.numpy()
    return out


def eval_case(ref, seed):
    """The reference's own eval_sampled (scripts/train_gat_custom.py:184-210) on a small model, fixed numpy seed."""
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    nu, ni, hidden = 30, 50, 128
    train_pos = {u: rng.integers(0, ni, size=int(rng.integers(1, 6))) for u in range(nu)}
    eval_pos = {int(u): int(rng.integers(0, ni)) for u in rng.permutation(nu)[:20]}
    ei = ref.build_edge_index(nu, ni, train_pos)
    feats = torch.nn.functional.normalize(torch.randn(ni, 128), dim=1)
    model = ref.CustomGAT(nu, ni, 128, hidden, 2).eval()
    with torch.no_grad():
        model.user_emb.weight.mul_(5.0)        # spread the scores so that ranks are not all ties
    cfg = ref.Config("p", "r", "s", "g", "e", "m", eval_neg_k=25)
    np.random.seed(1234)
    metrics = ref.eval_sampled(model, cfg, feats, ei, train_pos, eval_pos)
    out = {"n_users": np.int64(nu), "n_items": np.int64(ni), "neg_k": np.int64(25), "np_seed": np.int64(1234),
           "edge_index": ei.numpy(), "item_feats": feats.numpy(),
           "train_users": np.concatenate([np.full(len(v), k) for k, v in train_pos.items()]),
           "train_items": np.concatenate(list(train_pos.values())),
           "eval_users": np.array(list(eval_pos.keys())), "eval_items": np.array(list(eval_pos.values()))}
    for k, v in model.state_dict().items():
        out["param:" + k] = v.numpy()
    for k, v in metrics.items():
        out["metric:" + k] = np.float64(v)
    with torch.no_grad():
        out["z"] = model(feats, ei).numpy()
    return out


def knn_case(seed, n, dim, k, min_similarity, batch_size, n_centers):
    """Run the reference's graphs/build_ii_knn.py AS A SCRIPT (its code lives inside main()): the GCS client is replaced
    by a stand-in whose download hands over a local .npy and whose upload is a no-op; the script's own ``tmp/<name>.npz``
    is then read back.  Clustered, un-normalised rows so that the min_similarity filter both keeps and drops edges."""
    import runpy
    import shutil
    import tempfile
    from scipy.sparse import load_npz

    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((n_centers, dim))
    emb = centers[rng.integers(0, len(centers), n)] + 0.8 * rng.standard_normal((n, dim))
    emb = (emb * rng.uniform(0.5, 3.0, size=(n, 1))).astype(np.float32)
    work = tempfile.mkdtemp(prefix="b200gat_knn_")
    src = os.path.join(work, "emb.npy")
    np.save(src, emb)

    class _Blob:
        def download_to_filename(self, dst):
            shutil.copyfile(src, dst)

        def upload_from_filename(self, path):
            pass

    class _Bucket:
        def blob(self, path):
            return _Blob()

    class _Client:
        def __init__(self, project=None):
            pass

        def bucket(self, name):
            return _Bucket()

    for name in ("google", "google.cloud", "google.cloud.storage"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["google.cloud.storage"].Client = _Client
    sys.modules["google"].cloud = sys.modules["google.cloud"]
    sys.modules["google.cloud"].storage = sys.modules["google.cloud.storage"]
    argv, cwd = sys.argv, os.getcwd()
    try:
        os.chdir(work)
        sys.argv = ["build_ii_knn.py", "--project-id", "p", "--embeddings-path", "gs://b/emb.npy", "--output-prefix", "gs://b/out",
                    "--output-name", "ii", "--k", str(k), "--min-similarity", str(min_similarity), "--batch-size", str(batch_size)]
        runpy.run_path(os.path.join(REF, "graphs", "build_ii_knn.py"), run_name="__main__")
        m = load_npz(os.path.join(work, "tmp", "ii.npz"))
    finally:
        sys.argv = argv
        os.chdir(cwd)
        sys.modules["google.cloud.storage"].Client = object
    out = {"embeddings": emb, "k": np.int64(k), "min_similarity": np.float64(min_similarity), "batch_size": np.int64(batch_size),
           "rows": m.row.astype(np.int32), "cols": m.col.astype(np.int32), "sims": m.data.astype(np.float32)}
    shutil.rmtree(work, ignore_errors=True)
    return out


def main_knn():
    os.makedirs(OUT, exist_ok=True)
    # 128-d (fused features) with several similarity batches and a ragged last one; 384-d (text embeddings)
    np.savez_compressed(os.path.join(OUT, "knn_128.npz"), **knn_case(11, 333, 128, 20, 0.3, 100, 14))
    np.savez_compressed(os.path.join(OUT, "knn_384.npz"), **knn_case(12, 150, 384, 10, 0.25, 1000, 12))


def main():
    if sys.argv[1:] == ["knn"]:            # only the kNN fixtures (leaves the others untouched)
        main_knn()
        return
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    main_knn()
    np.savez_compressed(os.path.join(OUT, "custom_layer_plain.npz"), **layer_case(ref, 1, 160, 1500, 128, 1.0))
    np.savez_compressed(os.path.join(OUT, "custom_layer_clamped.npz"), **layer_case(ref, 2, 120, 1200, 128, 12.0, hub=7))
    np.savez_compressed(os.path.join(OUT, "custom_model.npz"),
                        **model_case(ref, 3, n_users=40, n_items=70, feat_dim=128, hidden=128, layers=2,
                                     n_inter=300, n_triples=256))
    np.savez_compressed(os.path.join(OUT, "eval_sampled.npz"), **eval_case(ref, 4))
    # edge-list builder: dict order, array order, duplicates
    tp = {3: np.array([5, 1, 5]), 0: np.array([2]), 7: np.array([0, 9, 9, 4])}
    np.savez_compressed(os.path.join(OUT, "edge_index_small.npz"), edge_index=ref.build_edge_index(8, 10, tp).numpy(),
                        keys=np.array(list(tp.keys())), lens=np.array([len(v) for v in tp.values()]),
                        items=np.concatenate(list(tp.values())))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
