"""CPU: the C-ABI library loads and exports every symbol the header declares; host-side contract checks
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200gat.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200gat_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import b200gat  # noqa: F401  (builds nothing; fails loudly if the .so is missing)
    from b200gat import _lib
    syms = _declared_symbols()
    assert len(syms) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200gat.h but not exported"
    assert sorted(_lib.EXPORTS) == syms, "ctypes binding and header disagree"
    assert lib.b200gat_abi_version() == 1


def test_argument_errors_come_back_as_messages():
    from b200gat import _lib
    with pytest.raises(RuntimeError, match="2\\^31"):
        _lib.graph_workspace_bytes(10, 2 ** 31)
    assert _lib.graph_workspace_bytes(100, 1000) > 8 * 4000
    assert _lib.loss_workspace_bytes(1000, 256) > 0


def test_cpu_tensors_are_rejected_not_silently_computed():
    import b200gat
    layer = b200gat.SimpleGATLayer(128, 128).eval()
    x = torch.randn(10, 128)
    ei = torch.randint(0, 10, (2, 30))
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(x, ei)
    conv = b200gat.GATConv(128, 128, heads=2, concat=False, add_self_loops=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        conv(x, ei)
    with pytest.raises(RuntimeError, match="CUDA"):
        b200gat.bpr_loss(x, 4, torch.zeros(3, dtype=torch.long), torch.zeros(3, dtype=torch.long), torch.zeros(3, dtype=torch.long))


def test_gatconv_rejects_unsupported_configurations():
    import b200gat
    for kw in (dict(), dict(concat=False), dict(concat=False, add_self_loops=False, edge_dim=4),
               dict(concat=False, add_self_loops=False, residual=True)):
        with pytest.raises(NotImplementedError):
            b200gat.GATConv(128, 128, **kw)
    conv = b200gat.GATConv(128, 128, heads=4, concat=False, add_self_loops=False, dropout=0.1)
    with pytest.raises(NotImplementedError):
        conv(torch.zeros(2, 128), torch.zeros(2, 1, dtype=torch.long), edge_attr=torch.zeros(1, 4))


def test_state_dict_keys_and_inits_match_the_reference_contract(golden_dir):
    import b200gat
    torch.manual_seed(0)
    m = b200gat.CustomGAT(40, 70, 128, 128, 2)
    g = np.load(os.path.join(golden_dir, "custom_model.npz"))
    ref_keys = sorted(k[len("param:"):] for k in g.files if k.startswith("param:"))
    assert sorted(m.state_dict().keys()) == ref_keys
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == g["param:" + k].shape, k
    m.load_state_dict({k: torch.from_numpy(g["param:" + k]).float() for k in ref_keys})
    lay = m.layers[0]
    assert lay.lin.weight.abs().max() <= (6 / 256) ** 0.5 + 1e-6
    big = b200gat.SimpleGATLayer(128, 128)
    assert big.a_src.abs().max() <= (6 / 129) ** 0.5 + 1e-6 and big.a_src.abs().max() > 0.15
    p = b200gat.PyGGAT(40, 70, 128, 128, 2, heads=4, attn_dropout=0.1)
    keys = set(p.state_dict().keys())
    assert {"user_emb.weight", "item_proj.weight", "item_proj.bias", "convs.0.att_src", "convs.0.att_dst", "convs.0.bias",
            "convs.0.lin.weight", "convs.1.lin.weight"} <= keys
    assert p.convs[0].att_src.shape == (1, 4, 128) and p.convs[0].lin.weight.shape == (512, 128)
    assert torch.count_nonzero(p.convs[0].bias) == 0
    # PyG <= 2.4 checkpoints name the shared projection lin_src / lin_dst
    sd = p.state_dict()
    for l in (0, 1):
        w = sd.pop(f"convs.{l}.lin.weight")
        sd[f"convs.{l}.lin_src.weight"] = w
        sd[f"convs.{l}.lin_dst.weight"] = w
    p.load_state_dict(sd)


def test_build_edge_index_matches_golden(golden_dir):
    import b200gat
    g = np.load(os.path.join(golden_dir, "edge_index_small.npz"))
    tp, off = {}, 0
    for k, n in zip(g["keys"], g["lens"]):
        tp[int(k)] = g["items"][off:off + n]
        off += n
    np.testing.assert_array_equal(b200gat.build_edge_index(8, 10, tp).numpy(), g["edge_index"])
    assert b200gat.build_edge_index(3, 3, {}).shape == (2, 0)


def test_synthetic_graph_shape():
    from b200gat import synth
    ei, feats = synth.make_graph(*synth.CONFIGS["tiny"])
    nu, ni, n_inter, k = synth.CONFIGS["tiny"]
    assert ei.shape == (2, 2 * n_inter + k * ni) and feats.shape == (ni, 128)
    ui = ei[:, :2 * n_inter]
    assert torch.equal(ui[0, 0::2], ui[1, 1::2]) and torch.equal(ui[1, 0::2], ui[0, 1::2])   # interleaved symmetric
    assert (ui[0, 0::2] < nu).all() and (ui[1, 0::2] >= nu).all()
    ii = ei[:, 2 * n_inter:]
    assert (ii >= nu).all() and (ii[0] != ii[1]).all()
    nb = ii[1].view(ni, k)
    assert all(len(set(r.tolist())) == k for r in nb)
    np.testing.assert_allclose(feats.norm(dim=1).numpy(), 1.0, rtol=1e-5)


def test_union_edge_index_appends_the_knn_block_in_coo_order(golden_dir):
    """U-I block first, then n_users + row -> n_users + col for every kNN COO entry, nothing sorted or de-duplicated
    (SURVEY.md 8d); the kNN COO is the reference script's own output (tests/golden/knn_128.npz)."""
    import b200gat
    g = np.load(os.path.join(golden_dir, "knn_128.npz"))
    rows, cols = torch.from_numpy(g["rows"]), torch.from_numpy(g["cols"])
    n_items = int(g["embeddings"].shape[0])
    nu = 7
    tp = {u: np.array([(3 * u + j) % n_items for j in range(u % 3 + 1)]) for u in range(nu)}
    ui = b200gat.build_edge_index(nu, n_items, tp)
    ei = b200gat.union_edge_index(ui, nu, rows, cols)
    assert ei.dtype == torch.int64 and ei.shape == (2, ui.shape[1] + rows.numel())
    assert torch.equal(ei[:, :ui.shape[1]], ui)
    assert torch.equal(ei[0, ui.shape[1]:], rows.long() + nu) and torch.equal(ei[1, ui.shape[1]:], cols.long() + nu)
    # same edges as the oracle's CSR sees them: in-degree of item i = #interactions + #times it is somebody's neighbour
    deg = torch.bincount(ei[1], minlength=nu + n_items)
    expect = torch.bincount(cols.long() + nu, minlength=nu + n_items) + torch.bincount(ui[1], minlength=nu + n_items)
    assert torch.equal(deg, expect)
    with pytest.raises(ValueError):
        b200gat.union_edge_index(ui.t(), nu, rows, cols)
    with pytest.raises(ValueError):
        b200gat.union_edge_index(ui, nu, rows, cols[:-1])


def test_ranking_metrics_host_logic_matches_reference_metrics(golden_dir):
    """b200gat.ranking_metrics is plain torch: on the ranks of the reference's own eval_sampled run (golden fixture, same
    numpy seed) it must return the reference's Recall/NDCG; empty input gives zeros (scripts/train_gat_custom.py:206-210)."""
    import b200gat
    from oracle import gat_oracle as O
    g = dict(np.load(os.path.join(golden_dir, "eval_sampled.npz")))
    nu, ni, neg_k = int(g["n_users"]), int(g["n_items"]), int(g["neg_k"])
    train_pos = {}
    for u, it in zip(g["train_users"], g["train_items"]):
        train_pos.setdefault(int(u), []).append(int(it))
    train_pos = {u: np.array(v) for u, v in train_pos.items()}
    eval_pos = {int(u): int(i) for u, i in zip(g["eval_users"], g["eval_items"])}
    np.random.seed(int(g["np_seed"]))
    users, cands = O.sample_eval_candidates(train_pos, eval_pos, ni, neg_k)
    ranks, _ = O.eval_ranks(torch.from_numpy(g["z"]), nu, users, cands)
    m = b200gat.ranking_metrics(ranks.to(torch.int32))
    for k in ("recall@10", "recall@20", "ndcg@10", "ndcg@20"):
        assert abs(m[k] - float(g["metric:" + k])) < 1e-12, (k, m[k])
    assert b200gat.ranking_metrics(torch.zeros(0, dtype=torch.int32)) == {"recall@10": 0.0, "recall@20": 0.0, "ndcg@10": 0.0,
                                                                          "ndcg@20": 0.0}
