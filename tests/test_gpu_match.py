"""GPU parity tests for brute-force descriptor matching: indices, accept flags and the per-row
best / second-best distances are BIT-EXACT against the oracle (same float32 evaluation order)."""
import numpy as np
import pytest

import replay
import synth
from backends import product

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    vo = product()
    c = vo.Context(0)
    yield c
    c.close()


def _rows_dev(ctx, A, B, row_begin=0, row_end=None):
    import torch
    n1, dim = A.shape
    row_end = n1 if row_end is None else row_end
    rows = row_end - row_begin
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    best = torch.empty(max(rows, 1), dtype=torch.float32, device="cuda")
    second = torch.empty_like(best)
    idx = torch.empty(max(rows, 1), dtype=torch.int32, device="cuda")
    pairs = torch.empty((max(rows, 1), 2), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    n, stats = ctx.match_dev(dA.data_ptr(), n1, dB.data_ptr(), len(B), dim, pairs.data_ptr(), rows, row_begin=row_begin,
                             row_end=row_end, d_best=best.data_ptr(), d_second=second.data_ptr(), d_idx=idx.data_ptr())
    torch.cuda.synchronize()
    return pairs[:n].cpu().numpy(), best[:rows].cpu().numpy(), second[:rows].cpu().numpy(), idx[:rows].cpu().numpy()


@pytest.mark.parametrize("n1,n2,dim", [(1, 1, 10), (1, 2, 10), (7, 3, 10), (127, 490, 10), (300, 1000, 10),
                                       (2049, 777, 10), (513, 4099, 10), (64, 64, 3), (100, 200, 4), (100, 200, 7),
                                       (100, 200, 8), (100, 200, 12), (100, 200, 16)])
def test_match_bit_exact(ctx, oracle, n1, n2, dim):
    A, B = synth.descriptors(n1, n2, dim=dim, seed=n1 * 31 + n2, dup_frac=0.02)
    idA = np.arange(n1, dtype=np.int32) % 50
    idB = np.arange(n2, dtype=np.int32) % 50
    pairs, stats = ctx.match(A, B, 0.2, 0.8, idA, idB)
    rp, rstats, rbest, rsecond, ridx = oracle.match(A, B, 0.2, 0.8, idA, idB, want_rows=True)
    assert np.array_equal(pairs, rp)
    assert stats == rstats
    p2, best, second, idx = _rows_dev(ctx, A, B)
    assert np.array_equal(p2, rp)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(best.view(np.uint32), rbest.view(np.uint32))
    assert np.array_equal(second.view(np.uint32), rsecond.view(np.uint32))


@pytest.mark.parametrize("n1,n2,noise,dup,scale", [
    (8192, 8192, 0.0, 0.02, 1.0), (9001, 12345, 0.1, 0.0, 1.0), (20000, 8200, 0.14, 0.3, 1.0),
    (40000, 9000, 0.0, 0.05, 1.0),    # two warps per 32-row group
    (80000, 8192, 0.05, 0.1, 1.0),    # one warp per group (the large-problem configuration)
    (9000, 9000, 0.05, 0.1, 37.5),    # descriptors far from unit scale: the filter's error bound is relative
    (9000, 9000, 0.05, 0.1, 3e16),    # magnitudes outside the filter's analysis: exact scan selected on the device
    (9000, 9000, 0.05, 0.1, 1e-17),
])
def test_match_indexed_path_bit_exact(ctx, oracle, n1, n2, noise, dup, scale):
    """both sets >= 8192 rows: Morton-ordered rows AND columns, tile boxes, tile skipping, tensor-core lower-bound
    filter + exact evaluation of the survivors, out-of-order column visits with the explicit lowest-index tie-break
    (many duplicate columns) - still bit-exact, values and indices of every row"""
    A, B = synth.descriptors(n1, n2, seed=n1 + n2, copy_frac=0.8, dup_frac=dup, noise=noise)
    A, B = (A * np.float32(scale)).astype(np.float32), (B * np.float32(scale)).astype(np.float32)
    rp, rstats, rbest, rsecond, ridx = oracle.match(A, B, want_rows=True, n_threads=8)
    p2, best, second, idx = _rows_dev(ctx, A, B)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(best.view(np.uint32), rbest.view(np.uint32))
    assert np.array_equal(second.view(np.uint32), rsecond.view(np.uint32))
    assert np.array_equal(p2, rp)
    # a row shard takes the same path and concatenates
    lo = max(0, min(n1 // 3, n1 - 8500))
    hi = min(n1, lo + 8500)
    ps, _, _, _ = _rows_dev(ctx, A, B, lo, hi)
    assert np.array_equal(ps, rp[(rp[:, 0] >= lo) & (rp[:, 0] < hi)])


@pytest.mark.parametrize("seed", [1, 2])
def test_match_indexed_path_adversarial_sweep(ctx, oracle, seed):
    """random sizes x {uniform, clustered, lattice, low-rank, heavy-tailed} x common scale 1e-3..1e8 x common offset
    up to 100x the spread x two threshold pairs: the cases where a lower-bound filter is most likely to be wrong
    (near-duplicates with aligned rounding errors, norms far above the distances, huge dynamic range). Seed 1 is the
    sweep that caught the filter's first, too small, error constant."""
    rng = np.random.default_rng(seed)
    for case in range(10):
        kind, A, B, dist_thr, ratio_thr = synth.stress_case(rng)
        rp, _, rbest, rsecond, ridx = oracle.match(A, B, dist_thr=dist_thr, ratio_thr=ratio_thr, want_rows=True, n_threads=16)
        import torch
        n1 = len(A)
        dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        best = torch.empty(n1, dtype=torch.float32, device="cuda"); second = torch.empty_like(best)
        idx = torch.empty(n1, dtype=torch.int32, device="cuda"); pairs = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
        n, _ = ctx.match_dev(dA.data_ptr(), n1, dB.data_ptr(), len(B), 10, pairs.data_ptr(), n1, dist_thr=dist_thr,
                             ratio_thr=ratio_thr, d_best=best.data_ptr(), d_second=second.data_ptr(), d_idx=idx.data_ptr())
        torch.cuda.synchronize()
        tag = f"seed {seed} case {case} {kind} {n1} x {len(B)}"
        assert np.array_equal(idx.cpu().numpy(), ridx), tag
        assert np.array_equal(best.cpu().numpy().view(np.uint32), rbest.view(np.uint32)), tag
        assert np.array_equal(second.cpu().numpy().view(np.uint32), rsecond.view(np.uint32)), tag
        assert np.array_equal(pairs[:n].cpu().numpy(), rp), tag


def test_match_few_rows_many_columns_takes_the_indexed_path(ctx, oracle):
    """300 query rows against 1,000,000 columns (a frame against a very large map): 10 row groups x 16 warps walk
    the column index; values, indices and pairs are still the reference's"""
    A, B = synth.descriptors(300, 1_000_000, seed=11, copy_frac=0.8, dup_frac=0.001, noise=0.02)
    rp, _, rbest, rsecond, ridx = oracle.match(A, B, want_rows=True, n_threads=8)
    p2, best, second, idx = _rows_dev(ctx, A, B)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(best.view(np.uint32), rbest.view(np.uint32))
    assert np.array_equal(second.view(np.uint32), rsecond.view(np.uint32))
    assert np.array_equal(p2, rp)


def test_match_indexed_path_offset_descriptors(ctx, oracle):
    """descriptors with a large common offset (|x| ~ 100, spread ~ 1): the filter centres them, the exact evaluation
    does not - results are those of the reference on the data as given"""
    A, B = synth.descriptors(9000, 10000, seed=3, copy_frac=0.8, dup_frac=0.01, noise=0.03)
    off = np.linspace(-120, 150, 10).astype(np.float32)
    A, B = (A + off).astype(np.float32), (B + off).astype(np.float32)
    rp, _, rbest, rsecond, ridx = oracle.match(A, B, want_rows=True, n_threads=8)
    p2, best, second, idx = _rows_dev(ctx, A, B)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(best.view(np.uint32), rbest.view(np.uint32))
    assert np.array_equal(second.view(np.uint32), rsecond.view(np.uint32))
    assert np.array_equal(p2, rp)


def test_match_indexed_path_descriptors_at_odd_float_offsets(ctx, oracle):
    """device pointers that are only 4-byte aligned (views into a larger float buffer at odd offsets, with a row
    offset on top): the index build's vectorised passes must not assume more than the C-ABI promises"""
    import torch
    A, B = synth.descriptors(9000, 10000, seed=5, copy_frac=0.8, dup_frac=0.01, noise=0.03)
    rp, _, rbest, rsecond, ridx = oracle.match(A, B, want_rows=True, n_threads=8)
    bufA = torch.zeros(A.size + 3, dtype=torch.float32, device="cuda")
    bufB = torch.zeros(B.size + 5, dtype=torch.float32, device="cuda")
    bufA[1:1 + A.size] = torch.from_numpy(A.ravel()).cuda()
    bufB[3:3 + B.size] = torch.from_numpy(B.ravel()).cuda()
    pA, pB = bufA.data_ptr() + 4, bufB.data_ptr() + 12
    assert pA % 8 == 4 and pB % 8 == 4
    for row_begin, row_end in ((0, len(A)), (301, len(A))):
        rows = row_end - row_begin
        best = torch.empty(rows, dtype=torch.float32, device="cuda")
        second = torch.empty_like(best)
        idx = torch.empty(rows, dtype=torch.int32, device="cuda")
        pairs = torch.empty((rows, 2), dtype=torch.int32, device="cuda")
        n, _ = ctx.match_dev(pA, len(A), pB, len(B), 10, pairs.data_ptr(), rows, row_begin=row_begin, row_end=row_end,
                             d_best=best.data_ptr(), d_second=second.data_ptr(), d_idx=idx.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(idx.cpu().numpy(), ridx[row_begin:row_end])
        assert np.array_equal(best.cpu().numpy().view(np.uint32), rbest[row_begin:row_end].view(np.uint32))
        assert np.array_equal(second.cpu().numpy().view(np.uint32), rsecond[row_begin:row_end].view(np.uint32))
        assert np.array_equal(pairs[:n].cpu().numpy(), rp[rp[:, 0] >= row_begin])


def test_match_indexed_path_nonfinite(ctx, oracle):
    """NaN / inf descriptors inside the indexed path: a NaN or inf distance never wins (`d < best` is false),
    on the filter exactly as in the reference loop"""
    A, B = synth.descriptors(9000, 9500, seed=77, copy_frac=0.8, dup_frac=0.01, noise=0.02)
    rng = np.random.default_rng(5)
    A[rng.integers(0, len(A), 40), rng.integers(0, 10, 40)] = np.nan
    B[rng.integers(0, len(B), 40), rng.integers(0, 10, 40)] = np.nan
    A[rng.integers(0, len(A), 10), rng.integers(0, 10, 10)] = np.inf
    B[rng.integers(0, len(B), 10), rng.integers(0, 10, 10)] = -np.inf
    rp, _, rbest, rsecond, ridx = oracle.match(A, B, want_rows=True, n_threads=8)
    p2, best, second, idx = _rows_dev(ctx, A, B)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(best.view(np.uint32), rbest.view(np.uint32))
    assert np.array_equal(second.view(np.uint32), rsecond.view(np.uint32))
    assert np.array_equal(p2, rp)


def test_match_noisy_and_threshold_edge(ctx, oracle):
    """noise puts many best distances near 0.2 and ratios near 0.8: exact rounding decides"""
    A, B = synth.descriptors(4000, 6000, seed=9, copy_frac=0.7, noise=0.14)
    pairs, _ = ctx.match(A, B)
    rp, _, rbest, rsecond, _ = oracle.match(A, B, want_rows=True)
    assert np.array_equal(pairs, rp)
    near = int((np.abs(rbest - 0.2) < 1e-3).sum() + (np.abs(rbest / rsecond - 0.8) < 1e-3).sum())
    assert near > 0  # the case is actually exercised


def test_match_ties_nan_and_single_candidate(ctx, oracle):
    rng = np.random.default_rng(1)
    B = rng.uniform(-1, 1, (40, 10)).astype(np.float32)
    B[7] = B[3]
    B[20] = B[3]          # three identical rows: lowest index wins, ratio 0/0 = NaN -> rejected
    A = B[[3, 5, 7, 11]].copy()
    A[1] += np.float32(1e-3)
    pairs, _ = ctx.match(A, B)
    rp, _, rbest, rsecond, ridx = oracle.match(A, B, want_rows=True)
    assert np.array_equal(pairs, rp)
    assert ridx[0] == 3 and 0 not in pairs[:, 0]  # NaN ratio rejects the exact duplicate
    # N2 = 1: second stays FLT_MAX -> accepted (my_utilities.h:104)
    p1, _ = ctx.match(B[:5], B[2:3])
    r1, _ = oracle.match(B[:5], B[2:3])
    assert np.array_equal(p1, r1) and len(p1) == 1
    # NaN / inf descriptors never win
    A2 = A.copy()
    A2[2, 4] = np.nan
    B2 = B.copy()
    B2[0, 0] = np.inf
    p3, _ = ctx.match(A2, B2)
    r3, _ = oracle.match(A2, B2)
    assert np.array_equal(p3, r3)


def test_match_row_sharding(ctx, oracle):
    """row blocks of A are independent: shards concatenate to the unsharded result"""
    A, B = synth.descriptors(3001, 2500, seed=4)
    full, fstats = ctx.match(A, B, idA=np.arange(3001, dtype=np.int32), idB=np.arange(2500, dtype=np.int32))
    parts, poss, corr = [], 0, 0
    bounds = [0, 700, 1500, 1501, 3001]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        p, st = ctx.match(A, B, idA=np.arange(3001, dtype=np.int32), idB=np.arange(2500, dtype=np.int32), row_begin=lo,
                          row_end=hi)
        parts.append(p)
        poss += st[0]
        corr += st[1]
        rp, _ = oracle.match(A, B, row_begin=lo, row_end=hi)
        assert np.array_equal(p, rp)
    assert np.array_equal(np.concatenate(parts), full)
    assert (poss, corr) == fstats


def test_match_empty(ctx):
    A, B = synth.descriptors(10, 10, seed=1)
    p, st = ctx.match(A, B, row_begin=4, row_end=4)
    assert len(p) == 0
    p, st = ctx.match(A, np.zeros((0, 10), np.float32))
    assert len(p) == 0


def test_match_dataset_kat(ctx, dataset):
    """exec/match_points_test.cpp as a KAT on the bundled data: every accepted pair has equal id_real"""
    for i in range(0, 120, 5):
        a, b = replay.frame(dataset, i), replay.frame(dataset, i + 1)
        pairs, (possible, correct) = ctx.match(a["desc"], b["desc"], 0.2, 0.8, a["id_real"], b["id_real"])
        assert correct == len(pairs) == possible
        assert np.array_equal(a["id_real"][pairs[:, 0]], b["id_real"][pairs[:, 1]])


def test_match_large_properties(ctx, oracle):
    """65536 x 1M (a BASELINE config-4 row shard): oracle on a row sample + structural properties"""
    import torch
    n1, n2 = 65536, 1 << 20
    A, B = synth.descriptors(n1, n2, seed=42)
    pairs, best, second, idx = _rows_dev(ctx, A, B)
    assert np.all(np.diff(pairs[:, 0]) > 0)
    assert np.all(best <= second)
    sample = np.random.default_rng(0).choice(n1, 48, replace=False)
    sample.sort()
    for r in sample:
        rp, _, rb, rs, ri = oracle.match(A, B, row_begin=int(r), row_end=int(r) + 1, want_rows=True, n_threads=8)
        assert ri[0] == idx[r] and rb[0] == best[r] and rs[0] == second[r]
        assert (len(rp) == 1) == bool((pairs[:, 0] == r).any())


def test_match_full_config4_all_rows(ctx, oracle):
    """BASELINE config 4 AT FULL SIZE, the workload bench.py times: 1,048,576 x 1,048,576, D = 10.  The indexed,
    tensor-core-filtered path (what AUTO runs) against the index-free exact scan (match_scan10_kernel: every column
    is a candidate, Eigen's order, itself oracle-tested above) on ALL rows - best, second best, index and the accepted
    pairs bit for bit - and against the CPU oracle on 4096 sampled rows."""
    n1 = n2 = 1 << 20
    A, B = synth.descriptors(n1, n2, seed=42)
    ctx.match_set_path(ctx.MATCH_INDEXED_FILTERED)
    pairs, best, second, idx = _rows_dev(ctx, A, B)
    ctx.match_set_path(ctx.MATCH_ORDERED)
    try:
        pairs_x, best_x, second_x, idx_x = _rows_dev(ctx, A, B)
    finally:
        ctx.match_set_path(ctx.MATCH_AUTO)
    assert np.array_equal(idx, idx_x)
    assert np.array_equal(best.view(np.uint32), best_x.view(np.uint32))
    assert np.array_equal(second.view(np.uint32), second_x.view(np.uint32))
    assert np.array_equal(pairs, pairs_x)
    assert 0.85 * n1 < len(pairs) < 0.95 * n1  # ~90 % of the rows are exact copies of a column
    sample = np.sort(np.random.default_rng(1).choice(n1, 4096, replace=False))
    rp, _, rb, rs, ri = oracle.match(np.ascontiguousarray(A[sample]), B, want_rows=True, n_threads=16)
    assert np.array_equal(ri, idx[sample])
    assert np.array_equal(rb.view(np.uint32), best[sample].view(np.uint32))
    assert np.array_equal(rs.view(np.uint32), second[sample].view(np.uint32))
    accepted = np.zeros(n1, bool)
    accepted[pairs[:, 0]] = True
    ref_acc = np.zeros(len(sample), bool)
    ref_acc[rp[:, 0]] = True
    assert np.array_equal(accepted[sample], ref_acc)


@pytest.mark.parametrize("path", [1, 2, 3, 4])
def test_match_paths_agree(ctx, oracle, path):
    """every selectable path (plain brute force, ordered exact scan, indexed exact, indexed + tensor-core filter)
    returns the oracle's (pairs, best, second, index) on a set large enough for the index"""
    A, B = synth.descriptors(9000, 20000, seed=77, noise=0.03, dup_frac=0.05)
    rp, _, rb, rs, ri = oracle.match(A, B, want_rows=True, n_threads=8)
    ctx.match_set_path(path)
    try:
        pairs, best, second, idx = _rows_dev(ctx, A, B)
    finally:
        ctx.match_set_path(0)
    assert np.array_equal(pairs, rp) and np.array_equal(idx, ri)
    assert np.array_equal(best.view(np.uint32), rb.view(np.uint32))
    assert np.array_equal(second.view(np.uint32), rs.view(np.uint32))


@pytest.mark.parametrize("seed", [1, 2])
def test_match_pruning_hostile_sets(ctx, oracle, seed):
    """Data the index cannot prune and the bf16 bound cannot separate (all columns in a few tight clusters, rows at a
    common distance from them): the filtered path falls back to evaluating whole tiles outright (the dense-tile guard of
    match_scan10_mma_kernel) and must still return the oracle's result bit for bit."""
    rng = np.random.default_rng(seed)
    n1, n2 = 9000, 30000
    cent = rng.uniform(-1, 1, (4, 10))
    B = (cent[rng.integers(0, 4, n2)] + rng.normal(0, 0.004, (n2, 10))).astype(np.float32)
    A = (cent[rng.integers(0, 4, n1)] + rng.normal(0, [0.15, 0.01][seed - 1], (n1, 10))).astype(np.float32)
    B[rng.integers(0, n2, 200)] = B[rng.integers(0, n2, 200)]  # exact duplicates inside the clusters: lowest index wins
    A[:500] = B[rng.integers(0, n2, 500)]
    rp, _, rb, rs, ri = oracle.match(A, B, 0.2, 0.8, want_rows=True, n_threads=8)
    pairs, best, second, idx = _rows_dev(ctx, A, B)
    assert np.array_equal(idx, ri)
    assert np.array_equal(best.view(np.uint32), rb.view(np.uint32))
    assert np.array_equal(second.view(np.uint32), rs.view(np.uint32))
    assert np.array_equal(pairs, rp)


@pytest.mark.parametrize("n1,n2,dim,n_shards", [(9000, 20000, 10, 4), (40000, 9000, 10, 8), (300, 1000, 7, 3), (5, 40, 10, 8)])
def test_match_sharded_equals_unsharded(ctx, oracle, n1, n2, dim, n_shards):
    """vo_match_sharded_dev shard by shard on one GPU (no communicator: the exchange is done here with an element-wise
    maximum, which is what the MAX all-reduce computes): every row is owned by exactly one shard - a Morton-order
    segment on the indexed path, an index block otherwise - and the merged result is the oracle's match list."""
    import torch
    A, B = synth.descriptors(n1, n2, dim=dim, seed=5 + n1, dup_frac=0.02, noise=0.02)
    rp, _ = oracle.match(A, B, 0.2, 0.8, n_threads=8)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    merged = torch.full((n1,), -1, dtype=torch.int32, device="cuda")
    owned_total = 0
    for shard in range(n_shards):
        m = torch.full((n1,), -7, dtype=torch.int32, device="cuda")
        ctx.match_sharded_dev(dA.data_ptr(), n1, dB.data_ptr(), n2, dim, shard, n_shards, m.data_ptr())
        torch.cuda.synchronize()
        assert int(m.min()) >= -1  # every element written
        assert not bool(((m >= 0) & (merged >= 0)).any())  # no row answered by two shards
        owned_total += int((m >= 0).sum())
        merged = torch.maximum(merged, m)
    pairs = torch.empty((n1, 2), dtype=torch.int32, device="cuda")
    n = ctx.match_compact_dev(merged.data_ptr(), n1, pairs.data_ptr(), n1)
    assert n == len(rp) == owned_total
    assert np.array_equal(pairs[:n].cpu().numpy(), rp)
    # one shard = the plain call
    m1 = torch.empty((n1,), dtype=torch.int32, device="cuda")
    n_one = ctx.match_sharded_dev(dA.data_ptr(), n1, dB.data_ptr(), n2, dim, 0, 1, m1.data_ptr(), pairs.data_ptr(), n1)
    assert n_one == len(rp) and torch.equal(m1, merged)
