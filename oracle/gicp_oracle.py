"""CPU oracle for the GICP hot path.  TEST INFRASTRUCTURE ONLY.

This file is a float64 numpy/scipy restatement of the algorithm in the
reference's ``python-implementation/gicp.py`` (cited below as ``gicp.py:L``),
plus the 3-D generalisation that SURVEY.md section 8c specifies (the reference
itself is 2-D only).  Nothing in the product path (``generalized-icp_b200/``,
``gicp.py`` at the repo root) may import it: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs do, and there only as the checker or as the timed CPU arm.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference ``gicp.py`` in the dev container on seeded config-1 / config-2
inputs and stores its outputs; ``tests/test_oracle.py`` checks every function
below against those fixtures (k-NN index sets and 1-NN indices bit-exact,
covariances / weights / loss / gradient to 1e-9, end-to-end transform and
iteration count with the fidelity inner solver).  The 3-D branch has no
reference implementation to pin against (the reference has no 3-D code); it is
pinned only through its 2-D specialisation sharing every line of code with the
3-D one, and says so: "3-D parity = restatement only".

Third-party arithmetic on the path (not vendored by the reference, unpinned
there): scipy.spatial.KDTree.query (gicp.py:24,132), scipy.optimize.fmin_cg
(gicp.py:152), numpy.cov / numpy.linalg.eig / inv (gicp.py:12,13,145).  Versions
used to generate the fixtures are stamped into each golden file.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import fmin_cg
from scipy.spatial import KDTree

LAMBDA_TANGENT = 100.0  # gicp.py:5   epsilon
LAMBDA_NORMAL = 10.0    # gicp.py:11  epsilon*0.1
K_DEFAULT = 6           # gicp.py:24  k=6 including the query point


# --------------------------------------------------------------------------
# k-NN (gicp.py:21-25)
# --------------------------------------------------------------------------
def knn_bruteforce(points, k, radius, queries=None):
    """Canonical k-NN rule: float64 squared distance
    ``(dx*dx + dy*dy) [+ dz*dz]`` (that association), ascending, ties broken by
    the lowest point index (stable argsort), bound exclusive (``d < radius``,
    gicp.py:24 ``distance_upper_bound``).  Missing slots get index N and
    distance inf exactly like scipy (gicp.py:25 drops ``idx == N``).

    Returns (idx (M,k) int64, dist (M,k) float64)."""
    P = np.asarray(points, dtype=np.float64)
    Q = P if queries is None else np.asarray(queries, dtype=np.float64)
    n = len(P)
    idx = np.full((len(Q), k), n, dtype=np.int64)
    dist = np.full((len(Q), k), np.inf)
    if n == 0:
        return idx, dist
    step = max(1, int(4e6 // max(n, 1)))
    for s in range(0, len(Q), step):
        q = Q[s:s + step]
        d = q[:, None, :] - P[None, :, :]
        d2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]
        if P.shape[1] == 3:
            d2 = d2 + d[..., 2] * d[..., 2]
        order = np.argsort(d2, axis=1, kind="stable")[:, :k]
        dd = np.sqrt(np.take_along_axis(d2, order, axis=1))
        ok = dd < radius
        kk = order.shape[1]
        idx[s:s + step, :kk] = np.where(ok, order, n)
        dist[s:s + step, :kk] = np.where(ok, dd, np.inf)
    return idx, dist


def knn_kdtree(points, k, radius, queries=None, workers=-1):
    """Same contract through scipy's cKDTree, batched (what gicp.py:21-24 calls
    per point).  Equal to :func:`knn_bruteforce` on tie-free inputs."""
    P = np.asarray(points, dtype=np.float64)
    Q = P if queries is None else np.asarray(queries, dtype=np.float64)
    tree = KDTree(P)
    dist, idx = tree.query(Q, k=k, distance_upper_bound=radius, workers=workers)
    if k == 1:
        dist, idx = dist[:, None], idx[:, None]
    return idx.astype(np.int64), dist


# --------------------------------------------------------------------------
# covariance model (gicp.py:5-17, 19-35)
# --------------------------------------------------------------------------
def covariances_from_neighbors(points, idx, lam_t=LAMBDA_TANGENT, lam_n=LAMBDA_NORMAL):
    """Plane-to-plane regularised covariance of every point's neighbourhood.

    gicp.py:12  sample covariance, ddof=1, of the surviving neighbours;
    gicp.py:13-16 (2-D) v = eigenvector of the LARGEST eigenvalue,
                 C = [v v_perp] diag(100, 10) [v v_perp]^T = 10 I + 90 v v^T;
    3-D n = eigenvector of the SMALLEST eigenvalue,
                 C = lam_t I - (lam_t - lam_n) n n^T  (identical in 2-D);
    gicp.py:27,33-34 a point with <= 1 surviving neighbour gets the identity;
    gicp.py:31-32 non-finite input -> LinAlgError -> identity."""
    P = np.asarray(points, dtype=np.float64)
    n, dim = P.shape
    k = idx.shape[1]
    valid = idx < n
    cnt = valid.sum(axis=1)
    nb = P[np.where(valid, idx, 0)]                       # (n,k,dim)
    w = valid[..., None].astype(np.float64)
    safe = np.maximum(cnt, 1)[:, None]
    mean = (nb * w).sum(axis=1) / safe
    dev = (nb - mean[:, None, :]) * w
    cov = np.einsum("nki,nkj->nij", dev, dev) / np.maximum(cnt - 1, 1)[:, None, None]
    out = np.empty((n, dim, dim))
    good = (cnt > 1) & np.isfinite(cov).all(axis=(1, 2))
    covg = np.where(good[:, None, None], cov, np.eye(dim))
    evals, evecs = np.linalg.eigh(covg)                   # ascending
    if dim == 2:
        v = evecs[:, :, 1]                                # largest
        out[:] = lam_n * np.eye(2) + (lam_t - lam_n) * v[:, :, None] * v[:, None, :]
    else:
        nrm = evecs[:, :, 0]                              # smallest
        out[:] = lam_t * np.eye(3) - (lam_t - lam_n) * nrm[:, :, None] * nrm[:, None, :]
    out[~good] = np.eye(dim)
    return out


def compute_covariances(points, radius, k=K_DEFAULT, lam_t=LAMBDA_TANGENT,
                        lam_n=LAMBDA_NORMAL, method="kdtree"):
    """gicp.py:19-35 for a whole cloud.  Returns (cov (N,d,d), idx (N,k))."""
    P = np.asarray(points, dtype=np.float64)
    if len(P) == 0:
        return np.zeros((0, P.shape[1], P.shape[1])), np.zeros((0, k), np.int64)
    if method == "kdtree":
        idx, _ = knn_kdtree(P, k, radius)
    else:
        idx, _ = knn_bruteforce(P, k, radius)
    return covariances_from_neighbors(P, idx, lam_t, lam_n), idx


# --------------------------------------------------------------------------
# transforms (gicp.py:37-50, 176-177)
# --------------------------------------------------------------------------
def rot2(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s], [s, c]])


def rot3(w):
    """Rodrigues: rotation matrix of the rotation vector w (3-D branch)."""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K + 0.5 * K @ K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def offset_to_matrix(offset, dim):
    """gicp.py:42-50: (tx,ty,theta) -> 3x3; 3-D: (t, rotvec) -> 4x4."""
    T = np.eye(dim + 1)
    if dim == 2:
        T[:2, :2] = rot2(offset[2])
        T[:2, 2] = offset[:2]
    else:
        T[:3, :3] = rot3(offset[3:6])
        T[:3, 3] = offset[:3]
    return T


def apply_transformation(cloud, T):
    """gicp.py:176-177 (dimension taken from T)."""
    d = T.shape[0] - 1
    cloud = np.asarray(cloud, dtype=np.float64)
    return cloud[:, :d] @ T[:d, :d].T + T[:d, d]


# --------------------------------------------------------------------------
# correspondences and weights (gicp.py:123-145)
# --------------------------------------------------------------------------
def correspond(transformed, target, d_max, method="kdtree"):
    """gicp.py:132 unbounded 1-NN; gicp.py:136 ``d > d_max`` rejects (strict).
    Returns (idx (N,) int64 with -1 for rejected, dist (N,), always the true
    1-NN distance)."""
    if method == "kdtree":
        idx, dist = knn_kdtree(target, 1, np.inf, queries=transformed)
    else:
        idx, dist = knn_bruteforce(target, 1, np.inf, queries=transformed)
    idx, dist = idx[:, 0], dist[:, 0]
    return np.where(dist > d_max, -1, idx), dist


def weights(src_cov_k, tgt_cov, idx):
    """gicp.py:143-145: W_i = inv(C_src,k[i] + C_tgt[j]); rejected rows are 0
    (gicp.py:124,137)."""
    n, d, _ = src_cov_k.shape
    W = np.zeros((n, d, d))
    m = idx >= 0
    if m.any():
        W[m] = np.linalg.inv(src_cov_k[m] + tgt_cov[idx[m]])
    return W


def corresponding_points(target, idx):
    """gicp.py:123,140: matched target point, (0,..) for rejected rows."""
    target = np.asarray(target, dtype=np.float64)
    out = np.zeros((len(idx), target.shape[1]))
    m = idx >= 0
    out[m] = target[idx[m]]
    return out


# --------------------------------------------------------------------------
# objective (gicp.py:52-76)
# --------------------------------------------------------------------------
def _R(offset, dim):
    return rot2(offset[2]) if dim == 2 else rot3(offset[3:6])


def loss(offset, src, tgt, W):
    """gicp.py:52-58: sum_i r_i^T W_i r_i, r_i = q_i - R p_i - t, on the
    untransformed source with the absolute transform."""
    dim = src.shape[1]
    r = tgt - src @ _R(offset, dim).T - offset[:dim]
    Wr = np.einsum("nij,nj->ni", W, r)
    return float(np.sum(r * Wr))


def grad_loss(offset, src, tgt, W):
    """gicp.py:60-76 (2-D).  3-D: same chain rule with dR/dw_a evaluated by
    central differences of Rodrigues is avoided - we use the exact derivative
    of R(w) through the left-Jacobian-free formula dR/dw_a = d/de R(w + e e_a),
    computed analytically below."""
    dim = src.shape[1]
    R = _R(offset, dim)
    r = tgt - src @ R.T - offset[:dim]
    Wr = np.einsum("nij,nj->ni", W, r)
    g = np.zeros(len(offset))
    g[:dim] = -2.0 * Wr.sum(axis=0)
    M = -2.0 * (Wr.T @ src)                                # gicp.py:72
    if dim == 2:
        th = offset[2]
        dR = np.array([[-np.sin(th), -np.cos(th)], [np.cos(th), -np.sin(th)]])
        g[2] = np.sum(M * dR)
    else:
        for a, dR in enumerate(_drot3(offset[3:6])):
            g[3 + a] = np.sum(M * dR)
    return g


def _drot3(w):
    """Exact dR/dw_a, a = 0..2 (Gallego & Yezzi 2015, eq. III.7)."""
    w = np.asarray(w, dtype=np.float64)
    th2 = float(w @ w)
    R = rot3(w)
    gens = [np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], float),
            np.array([[0, 0, 1], [0, 0, 0], [-1, 0, 0]], float),
            np.array([[0, -1, 0], [1, 0, 0], [0, 0, 0]], float)]
    if th2 < 1e-16:
        return [G.copy() for G in gens]
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    out = []
    I = np.eye(3)
    for a in range(3):
        v = np.cross(w, (I - R)[:, a])
        Va = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
        out.append((w[a] * K + Va) @ R / th2)
    return out


# --------------------------------------------------------------------------
# inner minimisers
# --------------------------------------------------------------------------
def inner_cg(x0, src, tgt, W):
    """Fidelity mode: gicp.py:148-154, scipy fmin_cg with the same arguments.
    Returns (x, fopt, warnflag, func_calls)."""
    out = fmin_cg(f=lambda x: loss(x, src, tgt, W), x0=x0,
                  fprime=lambda x: grad_loss(x, src, tgt, W),
                  disp=False, full_output=True)
    return out[0], float(out[1]), int(out[4]), int(out[2])


def inner_newton(x0, src, tgt, W, max_iter=100, gtol=1e-11):
    """Well-defined target of the inner problem: the minimiser of the frozen
    objective, by damped Gauss-Newton on per-point residuals (float64), using a
    left perturbation R <- exp(eta) R, t <- t + tau.  Returns (x, f, 0, iters).
    Independent of the device's reduced-quadratic-form solver by construction."""
    dim = src.shape[1]
    x = np.array(x0, dtype=np.float64)
    R, t = _R(x, dim).copy(), x[:dim].copy()
    theta = float(x[2]) if dim == 2 else None

    def f_of(Rm, tv):
        r = tgt - src @ Rm.T - tv
        return float(np.einsum("ni,nij,nj->", r, W, r)), r

    f, r = f_of(R, t)
    lam = 1e-9
    it = 0
    nrot = 1 if dim == 2 else 3
    for it in range(1, max_iter + 1):
        Rp = src @ R.T                                    # rotated source
        # J_i = d r_i / d(tau, eta): -I for tau;  -[eta]x (R p) for eta
        if dim == 2:
            Jrot = -np.stack([-Rp[:, 1], Rp[:, 0]], axis=1)[:, :, None]   # (n,2,1)
        else:
            Z = np.zeros(len(Rp))
            # r = q - exp(eta) Rp - t,  exp(eta) Rp ~ Rp - [Rp]x eta  ->  dr/d eta = [Rp]x
            Jrot = np.stack([np.stack([Z, -Rp[:, 2], Rp[:, 1]], 1),
                             np.stack([Rp[:, 2], Z, -Rp[:, 0]], 1),
                             np.stack([-Rp[:, 1], Rp[:, 0], Z], 1)], 1)
        J = np.concatenate([-np.broadcast_to(np.eye(dim), (len(src), dim, dim)), Jrot], axis=2)
        WJ = np.einsum("nij,njk->nik", W, J)
        H = np.einsum("nji,njk->ik", J, WJ)
        g = np.einsum("nji,nj->i", WJ, r)                  # J^T W r  (grad = 2 g)
        if np.max(np.abs(2 * g)) <= gtol * max(1.0, abs(f)):
            break
        accepted = False
        for _ in range(40):
            try:
                step = -np.linalg.solve(H + lam * np.diag(np.diag(H)) + 1e-300 * np.eye(len(g)), g)
            except np.linalg.LinAlgError:
                lam = max(lam * 10, 1e-6)
                continue
            Rn = (rot2(step[2]) if dim == 2 else rot3(step[3:6])) @ R
            tn = t + step[:dim]
            fn, rn = f_of(Rn, tn)
            pred = -float(g @ step)                 # predicted decrease of the GN model (>0)
            if fn <= f or pred <= 1e-11 * abs(f):   # below what f can resolve: trust the GN model
                if dim == 2:
                    theta += step[2]
                R, t, f, r = Rn, tn, fn, rn
                lam = max(lam * 0.1, 1e-12)
                accepted = True
                break
            lam = max(lam * 10, 1e-9)
        if not accepted or np.max(np.abs(step)) < 1e-14:
            break
    if dim == 2:
        x = np.array([t[0], t[1], theta])
        f = loss(x, src, tgt, W)
    else:
        x = np.concatenate([t, rotvec_from_matrix(R)])
    return x, f, 0, it


def rotvec_from_matrix(R):
    """Inverse of rot3 (principal branch)."""
    c = (np.trace(R) - 1.0) / 2.0
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2.0
    s = np.linalg.norm(v)
    if s < 1e-14:
        return v
    th = np.arctan2(s, c)
    return v * (th / s)


# --------------------------------------------------------------------------
# the registration loop (gicp.py:78-174)
# --------------------------------------------------------------------------
def gicp_oracle(source_points, target_points, max_iterations=100, tolerance=1e-6,
                max_distance_correspondence=150, max_distance_nearest_neighbors=50,
                k=K_DEFAULT, lam_t=LAMBDA_TANGENT, lam_n=LAMBDA_NORMAL,
                inner="newton", recompute_src_cov=True, knn="kdtree", record=True, covariance_model=0):
    """Restatement of gicp.py:78-174, any dimension d in {2,3}.

    inner = "cg"     fidelity (fmin_cg, gicp.py:152)
            "newton" converged minimiser of the frozen inner objective
    recompute_src_cov: True  = gicp.py:120 (k-NN covariances of the transformed
                               cloud every iteration),
                       False = the R_k C_0 R_k^T shortcut the engine uses.
    covariance_model: 0 = plane-to-plane (GICP, the reference), 1 = point-to-point (C_src = 0, C_tgt = I),
                      2 = point-to-plane (C_src = 0, C_tgt estimated) - the slides' generalisation table
                      (presentation/main.typ:446-462); 1 and 2 imply the R C R^T shortcut (C_src = 0).
    Returns a dict; ``T``/``all_T``/... mirror the reference's 7-tuple
    (gicp.py:174) and ``trace`` holds the per-iteration stage data used for
    teacher-forced parity tests."""
    src = np.asarray(source_points, dtype=np.float64)
    tgt = np.asarray(target_points, dtype=np.float64)
    dim = src.shape[1]
    nparam = 3 if dim == 2 else 6
    tgt_cov, tgt_knn = compute_covariances(tgt, max_distance_nearest_neighbors, k, lam_t, lam_n, knn)
    T = np.eye(dim + 1)
    all_T = [T]
    offset = np.zeros(nparam)
    last = np.inf
    src_cov0, src_knn = compute_covariances(src, max_distance_nearest_neighbors, k, lam_t, lam_n, knn)
    if covariance_model in (1, 2):
        src_cov0 = np.zeros_like(src_cov0)
        recompute_src_cov = False
        if covariance_model == 1:
            tgt_cov = np.broadcast_to(np.eye(dim), tgt_cov.shape).copy()
    hw_src, hw_tgt, all_src_cov, trace = [], [], [], []
    converged_at = None
    for it in range(max_iterations):
        moved = apply_transformation(src, T)                                   # gicp.py:119
        if recompute_src_cov:
            src_cov, _ = compute_covariances(moved, max_distance_nearest_neighbors, k, lam_t, lam_n, knn)
        else:
            Rk = T[:dim, :dim]
            src_cov = Rk @ src_cov0 @ Rk.T
        all_src_cov.append(src_cov)
        idx, dist = correspond(moved, tgt, max_distance_correspondence, knn)    # gicp.py:129-138
        q = corresponding_points(tgt, idx)
        W = weights(src_cov, tgt_cov, idx)                                      # gicp.py:143-145
        if inner == "cg":
            x, fopt, warn, calls = inner_cg(offset, src, q, W)
        else:
            x, fopt, warn, calls = inner_newton(offset, src, q, W)
        offset = x
        delta = abs(last - fopt)                                                # gicp.py:155
        if record:
            trace.append(dict(T=T.copy(), idx=idx, dist=dist, W=W, q=q, offset=np.array(x),
                              min_loss=fopt, warnflag=warn, calls=calls))
        if delta < tolerance:                                                   # gicp.py:160
            converged_at = it
            break
        last = fopt
        T = offset_to_matrix(offset, dim)                                       # gicp.py:166
        all_T.append(T)
        order = np.argsort(np.linalg.det(W))[-5:]                               # gicp.py:170
        hw_src.append(moved[order])
        hw_tgt.append(q[order])
    return dict(T=T, all_T=all_T, src_cov0=src_cov0, tgt_cov=tgt_cov, hw_src=hw_src, hw_tgt=hw_tgt,
                all_src_cov=all_src_cov, converged_at=converged_at, n_outer=len(all_src_cov),
                trace=trace, src_knn=src_knn, tgt_knn=tgt_knn)
