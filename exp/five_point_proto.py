"""numpy prototype of cv::findEssentialMat(RANSAC) (OpenCV 4.x five-point.cpp + ptsetreg.cpp), checked against the
committed cv2 4.13 fixtures.  The C++ oracle / CUDA versions follow this file."""
import itertools, sys, os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# monomials of degree <= 3 in (x, y, z); Nister's elimination order
MONO = [(3,0,0),(0,3,0),(2,1,0),(1,2,0),(2,0,1),(2,0,0),(0,2,1),(0,2,0),(1,1,1),(1,1,0),
        (1,0,2),(1,0,1),(1,0,0),(0,1,2),(0,1,1),(0,1,0),(0,0,3),(0,0,2),(0,0,1),(0,0,0)]
MIDX = {m: i for i, m in enumerate(MONO)}

class RNG:
    def __init__(self, state=0xffffffffffffffff):
        self.state = state if state else 0xffffffff
    def next(self):
        self.state = ((self.state & 0xffffffff) * 4164903690 + (self.state >> 32)) & 0xffffffffffffffff
        return self.state & 0xffffffff
    def uniform(self, a, b):
        return a if a == b else self.next() % (b - a) + a

def pmul(p, q):
    r = {}
    for (a, ca) in p.items():
        for (b, cb) in q.items():
            m = (a[0]+b[0], a[1]+b[1], a[2]+b[2])
            r[m] = r.get(m, 0.0) + ca*cb
    return r
def padd(p, q, s=1.0):
    r = dict(p)
    for m, c in q.items():
        r[m] = r.get(m, 0.0) + s*c
    return r

def cv_null_basis(Q):
    """rows 5..8 of Vt of cv::SVD::compute(Q 5x9, MODIFY_A | FULL_UV) (lapack.cpp JacobiSVDImpl_: the rows of Q are
    orthogonalised by one-sided Jacobi, sorted by norm; the 4 missing rows are +-1/9 sign vectors from
    RNG(0x12345678), Gram-Schmidt'ed twice against all previous rows, normalised)."""
    m, n, n1 = 9, 5, 9
    At = np.zeros((9, 9)); At[:5] = Q
    eps = np.finfo(np.float64).eps * 10
    W = (At[:5]**2).sum(1)
    for it in range(max(m, 30)):
        changed = False
        for i in range(n-1):
            for j in range(i+1, n):
                a, b = W[i], W[j]
                p = float(At[i] @ At[j])
                if abs(p) <= eps*np.sqrt(a*b): continue
                p *= 2
                beta = a - b; gamma = np.hypot(p, beta)
                if beta < 0:
                    delta = (gamma - beta)*0.5
                    s_ = np.sqrt(delta/gamma); c_ = p/(gamma*s_*2)
                else:
                    c_ = np.sqrt((gamma + beta)/(gamma*2)); s_ = p/(gamma*c_*2)
                t0 = c_*At[i] + s_*At[j]; t1 = -s_*At[i] + c_*At[j]
                At[i], At[j] = t0, t1
                W[i], W[j] = (t0**2).sum(), (t1**2).sum()
                changed = True
        if not changed: break
    W = np.sqrt((At[:5]**2).sum(1))
    for i in range(n-1):
        j = i
        for k in range(i+1, n):
            if W[j] < W[k]: j = k
        if i != j:
            W[[i, j]] = W[[j, i]]; At[[i, j]] = At[[j, i]]
    rng = RNG(0x12345678)
    tiny = sys.float_info.min
    for i in range(n1):
        sd = W[i] if i < n else 0.0
        ii = 0
        while ii < 100 and sd <= tiny:
            val0 = 1.0/m
            for k in range(m):
                At[i, k] = val0 if (rng.next() & 256) != 0 else -val0
            for _ in range(2):
                for j in range(i):
                    sd = float(At[i] @ At[j])
                    At[i] = At[i] - sd*At[j]
                    asum = np.abs(At[i]).sum()
                    asum = 1/asum if asum > eps*100 else 0.0
                    At[i] *= asum
            sd = np.sqrt((At[i]**2).sum())
            ii += 1
        At[i] *= (1/sd if sd > tiny else 0.0)
    return At[5:9].copy()

def cv_solve_poly(c, max_iters=300):
    """cv::solvePoly (mathfuncs.cpp): Durand-Kerner, roots initialised to (1+i)^k, updated in place, 300 sweeps.
    c[k] = coefficient of z^k. Returns the roots in OpenCV's order."""
    n = len(c) - 1
    while n > 1 and abs(c[n]) <= np.finfo(np.float64).eps: n -= 1
    roots = []
    p = complex(1, 0); r = complex(1, 1)
    for i in range(n):
        roots.append(p); p = p*r
    for it in range(max_iters):
        max_diff = 0.0
        for i in range(n):
            p = roots[i]
            num = complex(c[n]); den = complex(c[n])
            for j in range(n):
                num = num*p + c[n-j-1]
                if j != i and p != roots[j]:
                    den = den*(p - roots[j])
            num = num/den
            roots[i] = p - num
            max_diff = max(max_diff, abs(num))
        if max_diff <= 0: break
    return roots

def five_point(q1, q2):
    """q1, q2: 5x2 normalised points. Returns list of 3x3 E (unit Frobenius)."""
    Q = np.stack([q2[:,0]*q1[:,0], q2[:,0]*q1[:,1], q2[:,0], q2[:,1]*q1[:,0], q2[:,1]*q1[:,1], q2[:,1], q1[:,0], q1[:,1], np.ones(5)], 1)
    EE = cv_null_basis(Q)  # 4 x 9 null-space basis (rows), OpenCV's
    # E(x,y,z) = x E0 + y E1 + z E2 + E3, entries are degree-1 polynomials
    Ep = [[{(1,0,0): EE[0,3*i+j], (0,1,0): EE[1,3*i+j], (0,0,1): EE[2,3*i+j], (0,0,0): EE[3,3*i+j]} for j in range(3)] for i in range(3)]
    # det E
    def det3(M):
        t = pmul(M[0][0], padd(pmul(M[1][1], M[2][2]), pmul(M[1][2], M[2][1]), -1))
        t = padd(t, pmul(M[0][1], padd(pmul(M[1][0], M[2][2]), pmul(M[1][2], M[2][0]), -1)), -1)
        t = padd(t, pmul(M[0][2], padd(pmul(M[1][0], M[2][1]), pmul(M[1][1], M[2][0]), -1)))
        return t
    cons = [det3(Ep)]
    # EEt
    EEt = [[None]*3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            s = {}
            for k in range(3): s = padd(s, pmul(Ep[i][k], Ep[j][k]))
            EEt[i][j] = s
    tr = padd(padd(EEt[0][0], EEt[1][1]), EEt[2][2])
    for i in range(3):
        for j in range(3):
            s = {}
            for k in range(3): s = padd(s, pmul(EEt[i][k], Ep[k][j]), 2.0)
            s = padd(s, pmul(tr, Ep[i][j]), -1.0)
            cons.append(s)
    A = np.zeros((10, 20))
    for r, p in enumerate(cons):
        for m, c in p.items(): A[r, MIDX[m]] = c
    A = np.linalg.solve(A[:, :10], A[:, 10:])  # reduced: rows = monomials 0..9 expressed by the last 10
    # rows 4..9: x^2 z, x^2, y^2 z, y^2, xyz, xy
    B = np.zeros((3, 13))
    for i in range(3):
        a1, a2 = A[2*i+4], A[2*i+5]
        r1 = np.zeros(13); r2 = np.zeros(13)
        r1[1:4] = a1[0:3]; r1[5:8] = a1[3:6]; r1[9:13] = a1[6:10]
        r2[0:3] = a2[0:3]; r2[4:7] = a2[3:6]; r2[8:12] = a2[6:10]
        B[i] = r1 - r2
    # det B(z): columns: cubic, cubic, quartic (coefficients highest power first)
    P = [[np.poly1d(B[j, 0:4]), np.poly1d(B[j, 4:8]), np.poly1d(B[j, 8:13])] for j in range(3)]
    det = (P[0][0]*(P[1][1]*P[2][2] - P[1][2]*P[2][1]) - P[0][1]*(P[1][0]*P[2][2] - P[1][2]*P[2][0])
           + P[0][2]*(P[1][0]*P[2][1] - P[1][1]*P[2][0]))
    cc = np.zeros(11); cc[:len(det.coeffs)] = det.coeffs[::-1]
    roots = cv_solve_poly(cc)
    Es = []
    for rt in roots:
        if abs(rt.imag) > 1e-10: continue
        z = rt.real
        Bz = np.array([[P[j][0](z), P[j][1](z), P[j][2](z)] for j in range(3)])
        _, _, vt = np.linalg.svd(Bz)
        xy1 = vt[2]
        if abs(xy1[2]) < 1e-10: continue
        x, y = xy1[0]/xy1[2], xy1[1]/xy1[2]
        Ev = EE[0]*x + EE[1]*y + EE[2]*z + EE[3]
        Ev = Ev/np.linalg.norm(Ev)
        Es.append(Ev.reshape(3,3))
    return Es

def sampson_err(E, q1, q2):
    x1 = np.concatenate([q1, np.ones((len(q1),1))], 1); x2 = np.concatenate([q2, np.ones((len(q2),1))], 1)
    Ex1 = x1 @ E.T; Etx2 = x2 @ E
    x2tEx1 = (x2*Ex1).sum(1)
    return (x2tEx1*x2tEx1/(Ex1[:,0]**2 + Ex1[:,1]**2 + Etx2[:,0]**2 + Etx2[:,1]**2)).astype(np.float32)

def update_niters(p, ep, model_points, max_iters):
    p = min(max(p, 0.), 1.); ep = min(max(ep, 0.), 1.)
    num = max(1. - p, sys.float_info.min)
    denom = 1. - (1. - ep)**model_points
    if denom < sys.float_info.min: return 0
    num = np.log(num); denom = np.log(denom)
    if denom >= 0 or -num >= max_iters*(-denom): return max_iters
    return int(np.rint(num/denom))  # cvRound: round half to even

def find_essential_ransac(x1, x2, K, prob=0.999, threshold=1.0, max_iters=1000):
    K = K.astype(np.float64)
    fx, fy, cx, cy = K[0,0], K[1,1], K[0,2], K[1,2]
    q1 = np.stack([(x1[:,0].astype(np.float64) - cx)/fx, (x1[:,1].astype(np.float64) - cy)/fy], 1)
    q2 = np.stack([(x2[:,0].astype(np.float64) - cx)/fx, (x2[:,1].astype(np.float64) - cy)/fy], 1)
    thr = threshold/((fx+fy)/2)
    t = np.float32(thr*thr)
    n = len(q1)
    rng = RNG()
    best, best_mask, max_good = None, None, 0
    niters = max(max_iters, 1)
    it = 0
    log = []
    while it < niters:
        idx = []
        for i in range(5):
            v = rng.uniform(0, n)
            while v in idx: v = rng.uniform(0, n)
            idx.append(v)
        Es = five_point(q1[idx], q2[idx])
        for E in Es:
            err = sampson_err(E, q1, q2)
            mask = err <= t
            good = int(mask.sum())
            if good > max(max_good, 4):
                best, best_mask, max_good = E, mask, good
                niters = update_niters(prob, (n - good)/n, 5, niters)
        log.append((idx, len(Es)))
        it += 1
    return best, best_mask, it, log

if __name__ == "__main__":
    fx = np.load(os.path.join(ROOT, "tests/golden/cv2_fixtures.npz"))
    rp = np.load(os.path.join(ROOT, "tests/golden/cv2_recoverpose.npz"))
    K = fx["K"]
    cases = [("ds%d" % i, fx) for i in range(7)] + [("syn%d" % i, fx) for i in range(5)] + [("c%d" % i, rp) for i in range(20)]
    for name, f in cases:
        x1, x2, Ecv = f[name+"_x1"], f[name+"_x2"], f[name+"_E"]
        if not np.any(Ecv): print(name, "cv2 returned nothing"); continue
        E, mask, iters, log = find_essential_ransac(x1, x2, K)
        if E is None: print(name, "no model"); continue
        d = min(np.abs(E - Ecv).max(), np.abs(E + Ecv).max())
        rm = f[name+"_ransac_mask"].ravel() if name+"_ransac_mask" in f else None
        print(f"{name}: n={len(x1)} iters={iters} inliers={int(mask.sum())} |E-Ecv|={d:.3e}", "mask_equal=%s" % (np.array_equal(mask, rm != 0) if rm is not None else "n/a"))
