"""Seeded synthetic workloads of the shapes BASELINE.json names.
Shared by the tests, __graft_entry__.smoke() and bench.py. numpy only."""
import numpy as np

K_REF = np.array([[180, 0, 320], [0, 180, 240], [0, 0, 1]], np.float32)
ROWS, COLS = 480, 640


def euler_pose(dx):
    """v2tEuler (reference src/defs.h:131-136) in float64 -> 3x4."""
    a, b, c = dx[3:6]
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    T = np.zeros((3, 4))
    T[:, :3] = Rx @ Ry @ Rz
    T[:, 3] = dx[:3]
    return T


def picp_frame(n=1 << 20, seed=42, permute=False, outlier_frac=0.10, invalid_frac=0.02, noise_px=0.5,
               n_world=None):
    """One PICP frame (config 2/3): world points generated in the GT camera frame
    (u~U(0,639), v~U(0,479), z~U(0.5,5)), back-projected, moved to the world by a GT pose;
    measurements = projection + N(0,noise_px); `outlier_frac` gross outliers (uniform pixel);
    `invalid_frac` points behind the camera; initial pose = perturbation o GT (|dt|=0.05, 0.02 rad
    per axis). Correspondences: identity (variant A) or a random permutation (variant B)."""
    rng = np.random.Generator(np.random.Philox(seed))
    n_world = n if n_world is None else n_world
    Kd = K_REF.astype(np.float64)
    u = rng.uniform(0, 639, n_world)
    v = rng.uniform(0, 479, n_world)
    z = rng.uniform(0.5, 5.0, n_world)
    cam = np.stack([(u - Kd[0, 2]) / Kd[0, 0] * z, (v - Kd[1, 2]) / Kd[1, 1] * z, z], 1)
    bad = rng.random(n_world) < invalid_frac
    cam[bad, 2] *= -1.0  # behind the camera
    gt = euler_pose(np.array([0.3, -0.2, 0.5, 0.05, -0.03, 0.08]))  # world-in-camera GT
    Rg, tg = gt[:, :3], gt[:, 3]
    world = (cam - tg) @ Rg  # R^T (c - t)
    meas = np.stack([u, v], 1) + rng.normal(0, noise_px, (n_world, 2))
    out = rng.random(n_world) < outlier_frac
    meas[out] = np.stack([rng.uniform(0, 639, out.sum()), rng.uniform(0, 479, out.sum())], 1)
    d = np.random.Generator(np.random.Philox(4242)).normal(size=3)  # same initial pose for every shard/seed
    d *= 0.05 / np.linalg.norm(d)
    pert = euler_pose(np.array([d[0], d[1], d[2], 0.02, -0.02, 0.02]))
    pose0 = np.zeros((3, 4))
    pose0[:, :3] = pert[:, :3] @ Rg
    pose0[:, 3] = pert[:, :3] @ tg + pert[:, 3]
    if permute:
        perm = rng.permutation(n_world).astype(np.int32)[:n]
        # image points stay in order (first ascending), world indices are scattered (second arbitrary)
        world_p = np.empty_like(world)
        world_p[perm] = world[:n] if n == n_world else world[perm]
        if n == n_world:
            world = world_p
            pairs = np.stack([np.arange(n, dtype=np.int32), perm], 1)
        else:
            pairs = np.stack([perm, perm], 1)
    else:
        pairs = np.stack([np.arange(n, dtype=np.int32), np.arange(n, dtype=np.int32)], 1)
    return dict(K=K_REF.copy(), rows=ROWS, cols=COLS, world=world.astype(np.float32),
                image=meas.astype(np.float32), pairs=np.ascontiguousarray(pairs, np.int32),
                pose0=pose0.astype(np.float32), pose_gt=gt.astype(np.float32))


def descriptors(n1, n2, dim=10, seed=42, copy_frac=0.9, dup_frac=0.001, noise=0.0):
    """Config 4: descB ~ U(-1,1)^dim with `dup_frac` exact duplicate rows; descA: `copy_frac` exact
    copies of random rows of descB (+ optional N(0,noise)), the rest fresh U(-1,1)^dim."""
    rng = np.random.Generator(np.random.Philox(seed))
    B = rng.uniform(-1, 1, (n2, dim)).astype(np.float32)
    ndup = int(n2 * dup_frac)
    if ndup and n2 > 1:
        src = rng.integers(0, n2, ndup)
        dst = rng.integers(0, n2, ndup)
        B[dst] = B[src]
    A = rng.uniform(-1, 1, (n1, dim)).astype(np.float32)
    if n2 > 0:
        cp = rng.random(n1) < copy_frac
        A[cp] = B[rng.integers(0, n2, int(cp.sum()))]
    if noise > 0:
        A = (A + rng.normal(0, noise, A.shape)).astype(np.float32)
    return np.ascontiguousarray(A), np.ascontiguousarray(B)


def stress_descriptors(rng, kind, n1, n2):
    """Adversarial descriptor sets for the matcher (exp/match_stress.py, tests/test_gpu_match.py): clustered, lattice
    (masses of exact ties and zero distances), low-rank, heavy-tailed; 60 % of the rows are (noisy) copies of
    columns; a random common scale (1e-3 .. 1e8) and a random common offset (up to 100x the spread)."""
    if kind == "uniform":
        B = rng.uniform(-1, 1, (n2, 10)); A = rng.uniform(-1, 1, (n1, 10))
    elif kind == "clustered":
        c = rng.uniform(-1, 1, (rng.integers(3, 40), 10))
        B = c[rng.integers(0, len(c), n2)] + rng.normal(0, 0.02, (n2, 10))
        A = c[rng.integers(0, len(c), n1)] + rng.normal(0, 0.02, (n1, 10))
    elif kind == "lattice":
        B = rng.integers(-2, 3, (n2, 10)) * 0.25; A = rng.integers(-2, 3, (n1, 10)) * 0.25
    elif kind == "lowrank":
        M = rng.normal(0, 1, (3, 10))
        B = rng.normal(0, 1, (n2, 3)) @ M; A = rng.normal(0, 1, (n1, 3)) @ M
    else:  # "cauchy"
        B = rng.standard_cauchy((n2, 10)); A = rng.standard_cauchy((n1, 10))
    A = A.astype(np.float32); B = B.astype(np.float32)
    cp = rng.random(n1) < 0.6
    A[cp] = B[rng.integers(0, n2, int(cp.sum()))] + rng.normal(0, rng.choice([0.0, 0.01, 0.1]), (int(cp.sum()), 10)).astype(np.float32)
    s = np.float32(rng.choice([1.0, 1.0, 1e-3, 250.0, 1e8]))
    off = (rng.uniform(-5, 5, 10) * rng.choice([0.0, 1.0, 100.0])).astype(np.float32)
    return np.ascontiguousarray(A * s + off * s), np.ascontiguousarray(B * s + off * s)


def stress_case(rng):
    """one random (kind, A, B, dist_thr, ratio_thr) sized to land on the matcher's indexed path"""
    kind = str(rng.choice(["uniform", "clustered", "lattice", "lowrank", "cauchy"]))
    n1 = int(rng.choice([33, 300, 4097, 9000, 20000])); n2 = int(rng.integers(8192, 40000))
    if n1 < 8192 and n1 * n2 < (1 << 28):
        n2 = max(n2, (1 << 28) // n1 + 1)
    n2 = min(n2, 1_200_000)
    A, B = stress_descriptors(rng, kind, n1, n2)
    if rng.random() < 0.5:
        thr = (0.2, 0.8)
    else:
        thr = (float(np.float32(np.median(np.abs(A)) ** 2 * 4 + 1e-30)), 1.5)
    return kind, A, B, thr[0], thr[1]


def picp_stress_case(rng):
    """Adversarial PICP frame (exp/picp_stress.py, tests/test_gpu_picp.py): random camera (pinhole or general K), any
    rotation, scene scale 1e-3..1e6, points in front of / behind / on the camera plane, border-pixel measurements,
    optional NaN / inf / overflowing points, random thresholds.  Returns
    (K, rows, cols, pose, world, image, pairs, thr, keep_outliers, general, scale)."""
    n = int(rng.choice([1, 3, 5, 1000, 1408, 1409, 50001, 300000]))
    scale = float(rng.choice([1e-3, 1.0, 1.0, 40.0, 1e6]))
    general = rng.random() < 0.4
    K = np.array([[rng.uniform(50, 900), 0, rng.uniform(100, 700)], [0, rng.uniform(50, 900), rng.uniform(100, 500)], [0, 0, 1]], np.float32)
    if general:
        K[0, 1] = rng.normal(0, 2); K[1, 0] = rng.normal(0, 0.5); K[2, 0] = rng.normal(0, 1e-4); K[2, 1] = rng.normal(0, 1e-4); K[2, 2] = rng.uniform(0.5, 2)
    rows, cols = int(rng.integers(100, 1000)), int(rng.integers(100, 1300))
    ang = rng.uniform(-3.1, 3.1, 3) * rng.choice([0.02, 1.0])
    pose = euler_pose(np.array([*(rng.normal(0, 1, 3) * scale), *ang])).astype(np.float32)
    # points: in front / behind / at z ~ 0 / far away / NaN / inf
    cam = np.stack([rng.normal(0, 1, n), rng.normal(0, 1, n), rng.uniform(-1, 6, n)], 1) * scale
    knd = rng.random(n)
    cam[knd < 0.03, 2] = rng.normal(0, 1e-6, int((knd < 0.03).sum())) * scale      # z ~ 0 both signs
    inject = rng.random() < 0.4   # non-finite / overflowing points make the reference's H non-finite too
    if inject:
        cam[(knd > 0.03) & (knd < 0.05)] *= 1e12                                       # overflow territory
    world = ((cam - pose[:, 3].astype(np.float64)) @ pose[:, :3].astype(np.float64)).astype(np.float32)
    if inject:
        world[(knd > 0.05) & (knd < 0.055)] = np.nan
        world[(knd > 0.055) & (knd < 0.06), int(rng.integers(0, 3))] = np.inf
    Kd = K.astype(np.float64)
    q = cam @ Kd.T
    with np.errstate(all="ignore"):
        uv = q[:, :2] / q[:, 2:3]
    image = (uv + rng.normal(0, rng.choice([0.0, 0.5, 30.0]), (n, 2))).astype(np.float32)
    image[~np.isfinite(image)] = 0
    edge = rng.random(n) < 0.05   # measurements exactly on the border pixels
    image[edge] = np.stack([rng.choice([0.0, cols - 1.0], int(edge.sum())), rng.choice([0.0, rows - 1.0], int(edge.sum()))], 1)
    pairs = np.stack([rng.integers(0, n, n), rng.integers(0, n, n)], 1).astype(np.int32)
    thr = float(rng.choice([1.0, 1000.0, 3000.0, 1e-6, 1e12]))
    keep = bool(rng.random() < 0.5)
    return K, rows, cols, pose, world, image, pairs, thr, keep, general, scale
