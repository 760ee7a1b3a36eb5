"""CPU: the invariants the kNN append pipeline (csrc/knn.cu, modes kModeGroupMax / kModeAppend + select_kernel) rests on, restated
in numpy on one row of similarities.  Nothing here runs device code: it checks that the BOUNDS the three sweeps hand to each other are
valid whatever the layout of the catalogue -- which is what makes the result independent of the sampling -- and that the ladder inside
an append sweep never drops a column of the final top-48.  Speed is the only thing the estimates may cost (tests/test_gpu_knn.py checks
the kernels themselves against the reference's output)."""
import numpy as np
import pytest

K_CAND, BLK, STRIDE_A, STRIDE_B = 48, 128, 16, 8     # kCand, kBlk, kStrideA, kStrideB of csrc/knn.cu


def _row(kind, n, rng):
    """One row of approximate similarities to n columns."""
    if kind == "random":
        return rng.standard_normal(n).astype(np.float32) * 0.09
    if kind == "clustered":              # 60 near neighbours spread over the catalogue
        s = rng.standard_normal(n).astype(np.float32) * 0.06
        s[rng.choice(n, 60, replace=False)] = 0.7 + 0.05 * rng.standard_normal(60).astype(np.float32)
        return s
    if kind == "sorted_in_sample":       # the row's cluster fills a SAMPLED block (block 16): thresholds come out high
        s = rng.standard_normal(n).astype(np.float32) * 0.06
        s[16 * BLK:17 * BLK] = 0.8 + 0.05 * rng.standard_normal(BLK).astype(np.float32)
        return s
    if kind == "sorted_off_sample":      # ... or blocks no sweep but the last one sees (blocks 1..7): thresholds come out low
        s = rng.standard_normal(n).astype(np.float32) * 0.06
        s[BLK:8 * BLK] = 0.8 + 0.05 * rng.standard_normal(7 * BLK).astype(np.float32)
        return s
    if kind == "ties":                   # many equal values: `>=` must re-admit the columns behind a threshold
        return np.round(rng.standard_normal(n) * 4).astype(np.float32) / 64
    raise ValueError(kind)


def _kth_largest(v, k):
    return np.sort(v)[-k]


def _group_max_threshold(s):
    """kModeGroupMax: every STRIDE_A-th block; 64 groups = (position in the 32-wide chunk) x (chunk parity); 48th largest maximum."""
    n_blocks = len(s) // BLK
    gmax = np.full(64, -np.inf, np.float32)
    for j in range(0, n_blocks, STRIDE_A):
        blk = s[j * BLK:(j + 1) * BLK].reshape(4, 32)               # 4 chunks of 32 columns
        for c in range(4):
            g = (c & 1) * 32
            gmax[g:g + 32] = np.maximum(gmax[g:g + 32], blk[c])
    top = np.sort(gmax)[::-1]
    return top[K_CAND - 1], max(top[K_CAND // 2 - 1] - top[K_CAND - 1], 1e-6)


def _append_sweep(s, thr0, dl, jstep, halves=2):
    """kModeAppend with the ladder.  Returns the appended column ids and the final threshold of every half."""
    n_blocks = len(s) // BLK
    appended, finals = [], []
    for half in range(halves):
        thr = np.float32(thr0)
        thr2, thr3 = np.float32(thr + dl), np.float32(thr + 2 * dl)
        cnt2 = cnt3 = 0
        for j in range(0, n_blocks, jstep):
            for c in range((4 // halves) * half, (4 // halves) * (half + 1)):
                col0 = j * BLK + c * 32
                v = s[col0:col0 + 32]
                g = v.reshape(4, 8).max(axis=1)                      # maxima of the four 8-column groups
                if not (g.max() >= thr):
                    continue                                         # (the kernel votes over 32 rows; for one row this is the same test)
                appended.extend(col0 + np.nonzero(v >= thr)[0])
                cnt2 += int((g >= thr2).sum())
                cnt3 += int((g >= thr3).sum())
                if cnt2 >= K_CAND:
                    thr, thr2, cnt2, thr3, cnt3 = thr2, thr3, cnt3, np.float32(thr3 + dl), 0
        finals.append(thr)
    return np.array(sorted(appended), dtype=np.int64), finals


@pytest.mark.parametrize("kind", ["random", "clustered", "sorted_in_sample", "sorted_off_sample", "ties"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_bounds_are_valid_for_any_layout(kind, seed):
    rng = np.random.default_rng(seed)
    n = 96 * BLK                                                      # >= kAppendMinBlocks column blocks
    s = _row(kind, n, rng)
    kth = _kth_largest(s, K_CAND)                                     # the row's 48th best similarity

    # sweep A: the 48th largest group maximum is attained by 48 different columns
    thr_a, dl_a = _group_max_threshold(s)
    assert thr_a <= kth
    assert (s >= thr_a).sum() >= K_CAND

    # sweep B: every STRIDE_B-th block re-admits those columns (sweep A's blocks are a subset), so its list holds >= 48 entries
    cols_b, fin_b = _append_sweep(s, thr_a, dl_a, STRIDE_B)
    assert len(cols_b) >= K_CAND
    thr_b = _kth_largest(s[cols_b], K_CAND)                           # select_kernel<false>
    assert thr_a <= thr_b <= kth
    dl_c = max(0.5 * (thr_b - thr_a), 1e-6)

    # sweep C (all blocks): the list holds >= 48 entries, contains every column of the true top-48 (up to ties at the 48th value),
    # and no column outside it beats the bound handed to the re-rank guard (the 48th best appended similarity)
    cols_c, fin_c = _append_sweep(s, thr_b, dl_c, 1)
    assert len(cols_c) >= K_CAND
    bound = _kth_largest(s[cols_c], K_CAND)                           # select_kernel<true>
    assert bound == kth
    assert max(fin_c) <= bound                                        # each half's final rung is backed by 48 appended columns
    outside = np.ones(n, bool)
    outside[cols_c] = False
    assert not outside.any() or s[outside].max() <= bound
    strictly_better = np.nonzero(s > kth)[0]
    assert np.isin(strictly_better, cols_c).all()


def test_ladder_reduces_appends_without_losing_candidates():
    """With a usable step the ladder appends fewer columns than the static threshold, and the top-48 is still complete."""
    rng = np.random.default_rng(7)
    n = 512 * BLK
    s = _row("random", n, rng)
    thr_a, dl_a = _group_max_threshold(s)
    cols_b, _ = _append_sweep(s, thr_a, dl_a, STRIDE_B)
    thr_b = _kth_largest(s[cols_b], K_CAND)
    static = int((s >= thr_b).sum())
    cols_c, _ = _append_sweep(s, thr_b, max(0.5 * (thr_b - thr_a), 1e-6), 1)
    assert len(cols_c) < static
    top = np.argsort(-s, kind="stable")[:K_CAND]
    assert np.isin(top[s[top] > _kth_largest(s, K_CAND)], cols_c).all()


def _fkey(f):
    """csrc/knn.cu:fkey -- order-preserving float32 -> uint32 (0 is below every real value)."""
    b = np.asarray(f, np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def _fkey_inv(k):
    k = np.asarray(k, np.uint32)
    return np.where(k & np.uint32(0x80000000), k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32).view(np.float32)


def test_select_kernel_bit_search_is_the_kth_largest():
    """select_kernel finds the 48th largest key bit by bit: the largest T with |{key >= T}| >= 48."""
    rng = np.random.default_rng(3)
    for trial in range(20):
        n = int(rng.integers(48, 1025))
        v = (rng.standard_normal(n) * rng.choice([1e-3, 0.1, 1.0])).astype(np.float32)
        if trial % 4 == 0:
            v = np.round(v * 8) / 8                               # ties, +0.0 and -0.0 among them
        keys = _fkey(v)
        assert np.array_equal(np.argsort(keys, kind="stable"), np.argsort(v + 0.0, kind="stable")) or np.all(np.diff(v[np.argsort(keys)]) >= 0)
        assert np.array_equal(_fkey_inv(keys).view(np.uint32), v.view(np.uint32))
        t = np.uint32(0)
        for bit in range(31, -1, -1):
            tr = t | np.uint32(1 << bit)
            if int((keys >= tr).sum()) >= K_CAND:
                t = tr
        assert _fkey_inv(t) == np.sort(v)[-K_CAND]
        assert int((keys > t).sum()) < K_CAND <= int((keys >= t).sum())
