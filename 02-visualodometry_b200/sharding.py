"""Host-side partitioning helpers for the multi-GPU paths.

PICP: contiguous blocks of the correspondence array per rank, one all-reduce of the 32-term
linearization per Gauss-Newton round (layout below = `result` of picp_linearize_kernel).
Matching: contiguous row blocks of the query set per rank, no collective; results concatenate in
rank order.  Pure index arithmetic, no device code."""
import numpy as np

N_TERMS = 32  # 21 upper-triangular H + 6 b + chi_in + chi_out + n_inliers + n_outliers + 1 pad


def shard_bounds(n, world_size):
    """[lo, hi) of every rank: contiguous, balanced to within one element, rank order = index order."""
    base, rem = divmod(int(n), int(world_size))
    bounds, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < rem else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def shard_range(n, world_size, rank):
    return shard_bounds(n, world_size)[rank]


def pack_terms(H, b, chi_in, chi_out, n_in, n_out):
    """6x6 H, 6 b and the stats -> the 32 doubles one rank contributes to the all-reduce."""
    out = np.zeros(N_TERMS, np.float64)
    H = np.asarray(H, np.float64)
    out[:21] = H[np.triu_indices(6)]
    out[21:27] = np.asarray(b, np.float64)
    out[27], out[28], out[29], out[30] = chi_in, chi_out, n_in, n_out
    return out


def unpack_terms(t):
    t = np.asarray(t, np.float64)
    H = np.zeros((6, 6))
    H[np.triu_indices(6)] = t[:21]
    H = H + np.triu(H, 1).T
    return dict(H=H, b=t[21:27].copy(), chi_in=float(t[27]), chi_out=float(t[28]), n_inliers=int(round(t[29])),
                n_outliers=int(round(t[30])))
