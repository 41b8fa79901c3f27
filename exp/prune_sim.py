"""CPU what-if: fraction of 128-column tiles a 32-row group must visit, for different curve / box choices
(bound = each row's true second-best distance, i.e. the floor no visiting order can beat)"""
import sys, os, numpy as np, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, "tests")); import synth
n = 1 << 20
A, B = synth.descriptors(n, n, seed=42)
rng = np.random.default_rng(0)

def morton(X, dims, bits):
    lo, hi = X[:, dims].min(0), X[:, dims].max(0)
    q = np.clip(((X[:, dims] - lo) / (hi - lo) * ((1 << bits) - 1)).astype(np.int64), 0, (1 << bits) - 1)
    key = np.zeros(len(X), np.int64)
    for b in range(bits):
        for i in range(len(dims)):
            key |= ((q[:, i] >> b) & 1) << (b * len(dims) + i)
    return key

def second_best(rows):
    out = np.empty(len(rows), np.float32)
    for i in range(0, len(rows), 64):
        a = A[rows[i:i + 64]]
        d = ((a * a).sum(1)[:, None] + (B * B).sum(1)[None, :] - 2 * a @ B.T)
        out[i:i + 64] = np.partition(d, 1, axis=1)[:, 1]
    return np.maximum(out, 0)

GROUP = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for dims, bits, boxdims in [((0, 4, 2, 6), 8, None)]:
    boxdims = boxdims or dims
    kb = morton(B, list(dims), bits); cb = np.argsort(kb, kind="stable")
    ka = morton(A, list(dims), bits); ra = np.argsort(ka, kind="stable")
    Bs = B[cb][:, boxdims].reshape(-1, 128, len(boxdims))
    blo, bhi = Bs.min(1), Bs.max(1)
    groups = rng.integers(0, n // GROUP, 48 * 32 // GROUP)
    fr = []
    for g in groups:
        rows = ra[g * GROUP:(g + 1) * GROUP]
        sb = second_best(rows)
        p = A[rows][:, boxdims]
        gap = np.maximum(0, np.maximum(blo[None] - p[:, None], p[:, None] - bhi[None]))
        need = ((gap ** 2).sum(2) < sb[:, None]).any(0)
        fr.append(need.mean())
    print(f"group {GROUP}: curve dims {dims} x {bits} bits, box dims {len(boxdims)}: tiles needed {100*np.mean(fr):.2f}% (median {100*np.median(fr):.2f}%)", flush=True)
