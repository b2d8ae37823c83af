"""Multi-path Pathfinder for many genes at once — what `BINARY pathfinder num_draws=1000
num_paths=C` does for ONE gene in /root/reference/bin/splotch_vinit:140, batched over genes x paths.

The algorithm is Stan's `pathfinder` method (Zhang, Carpenter, Gelman, Vehtari, JMLR 2022; CmdStan >= 2.33, not in
the reference tree — the reference only invokes it) restated from the paper with CmdStan's defaults
(history_size 5, num_elbo_draws 25, max_lbfgs_iters 1000, init_alpha 1e-3, tol_obj 1e-12, tol_rel_obj 1e4,
tol_grad 1e-8, tol_rel_grad 1e7, tol_param 1e-8, PSIS resampling over the paths):

  per path    L-BFGS on -lp from uniform(-2, 2); after every step the curvature pair (s, y) enters a history of
              `history_size` pairs and the diagonal estimate alpha is updated (Gilbert & Lemarechal); the BFGS
              inverse Hessian `diag(alpha) + B G B'` with `B = [alpha Y, S]` gives the normal approximation
              N(x + Sigma grad lp, Sigma) whose ELBO is estimated from `num_elbo_draws` draws; the iterate with the
              largest ELBO wins; `num_draws` draws from it with their lp and lp_approx.
  per gene    Pareto-smoothed importance resampling of the pooled draws of its paths (weights lp - lp_approx).

Every log-density / gradient is evaluated by the caller's `target` in batches of rows (one row = one gene-path or
one ELBO draw): for the product that is `Batch.log_prob_grad` on the GPU (`BatchTarget`); this module holds the host
control logic only and never evaluates a density itself.  Line search: weak-Wolfe bisection / expansion instead of
Stan's cubic-interpolating More-Thuente search (same acceptance conditions c1 = 1e-4, c2 = 0.9; the sequence of
iterates differs, the family of approximations does not).  The standard-normal variates of the ELBO and of the final
draws are common to all genes (keyed by seed, path and block, not by gene), and the starting points are keyed by
(seed, gene id, path), so a gene's result does not depend on the batch or shard it was processed in.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

EPS = np.finfo(np.float64).eps


# ------------------------------------------------------------------------------------------------ targets
class BatchTarget:
    """`api.Batch` as a Pathfinder target: rows are (theta, gene index) pairs evaluated on the GPU."""

    def __init__(self, batch):
        self.batch = batch
        self.dim = batch.model.dim
        self.G = batch.G
        self.gene_id0 = batch.gene_id0

    def lp_grad(self, theta, gene_of):
        return self.batch.log_prob_grad(theta, gene_of=gene_of, jacobian=True, want_grad=True)

    def lp(self, theta, gene_of):
        return self.batch.log_prob_grad(theta, gene_of=gene_of, jacobian=True, want_grad=False)


@dataclass
class PathfinderResult:
    draws: np.ndarray          # (G, n_out, D) unconstrained, PSIS-resampled over the paths of each gene
    lp: np.ndarray             # (G, n_out)  lp__        log density (jacobian adjusted) of each draw
    lp_approx: np.ndarray      # (G, n_out)  lp_approx__ log density of the normal approximation it came from
    path: np.ndarray           # (G, n_out)  path__      1-based path of origin
    pareto_k: np.ndarray       # (G,)
    elbo: np.ndarray           # (G, P) best ELBO of each path (-inf: the path failed)
    best_iter: np.ndarray      # (G, P) L-BFGS iteration of the best ELBO
    lbfgs_iters: np.ndarray    # (G, P)
    n_lp_grad: int = 0         # gradient evaluations spent
    n_lp: int = 0              # density-only evaluations spent
    info: dict = field(default_factory=dict)


# ------------------------------------------------------------------------------------------------ variates
def _normals(seed: int, purpose: int, path: int, block: int, shape, gene_id: int = 0):
    """Standard normals keyed by (seed, purpose, path, block, gene id).  The draws a gene returns (purpose 2) carry its
    id, so genes and paths get independent streams whatever the batch; the ELBO estimates (purpose 1) use common random
    numbers across genes on purpose (gene_id 0: one (num_elbo_draws, D) matrix per L-BFGS iteration instead of one
    per gene-path — each gene's estimate is unbiased either way, and the host generates 10^3 times fewer variates)."""
    key = (int(seed) & 0xFFFFFFFF) | (int(purpose) << 32) | (int(path) << 40)
    bg = np.random.Philox(key=key, counter=[int(block), int(gene_id), 0, 0])
    return np.random.Generator(bg).standard_normal(shape)


def _row_rng(seed: int, gene_id: int, path: int):
    key = (int(seed) & 0xFFFFFFFF) | (3 << 32) | (int(path) << 40)
    return np.random.Generator(np.random.Philox(key=key, counter=[int(gene_id), 0, 0, 0]))


# ------------------------------------------------------------------------------------------------ approximation
class _Approx:
    """Normal approximations N(mu, Sigma) of r rows, Sigma = diag(alpha) + B G B' (paper Alg. 4), in the factored form
    used for sampling (Alg. 5): with A = alpha^-1/2 B = Q R (thin QR) and M = I + R G R',
    Sigma^1/2 = alpha^1/2 (Q M^1/2 Q' + I - Q Q')."""

    def __init__(self, x, g, alpha, S, Y, m):
        # x, g, alpha: (r, D); g = gradient of -lp; S, Y: (r, J, D) oldest..newest in slots 0..m-1; m: (r,)
        r, D = x.shape
        self.r, self.D = r, D
        self.sqrt_alpha = np.sqrt(alpha)
        self.logdet = np.log(alpha).sum(axis=1)
        self.ok = np.isfinite(self.logdet)
        self.mu = x - alpha * g
        self.groups = []                         # (rows, Q (n, D, k), T (n, k, k) = M^1/2 - I), k = min(D, 2m)
        for mm in np.unique(m):
            rows = np.nonzero(m == mm)[0]
            if mm == 0:
                continue
            Sg, Yg = S[rows, :mm], Y[rows, :mm]                       # (n, mm, D)
            al = alpha[rows][:, None, :]
            E = np.triu(Sg @ np.swapaxes(Yg, 1, 2))                   # S'Y, upper triangle
            eta = np.einsum("nii->ni", E)
            with np.errstate(all="ignore"):
                try:
                    Einv = np.linalg.inv(E)
                except np.linalg.LinAlgError:
                    Einv = np.full_like(E, np.nan)
                YaY = (Yg * al) @ np.swapaxes(Yg, 1, 2)
                mid = YaY.copy()
                idx = np.arange(mm)
                mid[:, idx, idx] += eta
                G = np.zeros((len(rows), 2 * mm, 2 * mm))
                G[:, :mm, mm:] = -Einv
                G[:, mm:, :mm] = -np.swapaxes(Einv, 1, 2)
                G[:, mm:, mm:] = np.swapaxes(Einv, 1, 2) @ mid @ Einv
                B = np.concatenate([Yg * al, Sg], axis=1)                 # (n, 2mm, D) = B'
                # mean: x + Sigma grad lp = x - alpha g - B G B' g
                Bg = (B @ g[rows][:, :, None])[:, :, 0]
                self.mu[rows] -= ((G @ Bg[:, :, None])[:, None, :, 0] @ B)[:, 0]
                A = np.swapaxes(B / self.sqrt_alpha[rows][:, None, :], 1, 2)   # (n, D, 2mm)
                good = np.isfinite(A).all(axis=(1, 2)) & np.isfinite(G).all(axis=(1, 2))
                A[~good] = 0.0
                Q, R = np.linalg.qr(A)
                M = R @ np.where(good[:, None, None], G, 0.0) @ np.swapaxes(R, 1, 2)
                M = 0.5 * (M + np.swapaxes(M, 1, 2))
                kk = np.arange(M.shape[1])                 # min(D, 2mm): the dense case D < 2J of the paper
                M[:, kk, kk] += 1.0
                good &= np.isfinite(M).all(axis=(1, 2))
                M[~good] = np.eye(len(kk))
                w, V = np.linalg.eigh(M)
                good &= w.min(axis=1) > 0
                w = np.where(good[:, None], w, 1.0)
                T = (V * np.sqrt(w)[:, None, :]) @ np.swapaxes(V, 1, 2)
                T[:, kk, kk] -= 1.0
            self.logdet[rows] += np.log(w).sum(axis=1)
            self.ok[rows] &= good
            self.groups.append((rows, Q, T))
        self.ok &= np.isfinite(self.mu).all(axis=1)

    def draw(self, u):
        """u: (n, D) shared standard normals -> phi (r, n, D), log q (r, n)."""
        phi = np.broadcast_to(u[None], (self.r,) + u.shape).copy()
        for rows, Q, T in self.groups:
            Qu = u[None] @ Q                                            # (r, n, k), BLAS
            phi[rows] += (Qu @ np.swapaxes(T, 1, 2)) @ np.swapaxes(Q, 1, 2)
        phi *= self.sqrt_alpha[:, None, :]
        phi += self.mu[:, None, :]
        logq = -0.5 * (self.logdet[:, None] + (u * u).sum(axis=1)[None, :] + self.D * np.log(2 * np.pi))
        return phi, logq

    def take(self, rows):
        """Copy restricted to `rows` (indices into this object's rows)."""
        out = object.__new__(_Approx)
        out.r, out.D = len(rows), self.D
        out.sqrt_alpha, out.logdet, out.ok, out.mu = (self.sqrt_alpha[rows].copy(), self.logdet[rows].copy(),
                                                      self.ok[rows].copy(), self.mu[rows].copy())
        pos = -np.ones(self.r, dtype=np.int64)
        pos[rows] = np.arange(len(rows))
        out.groups = []
        for grows, Q, T in self.groups:
            sel = np.nonzero(pos[grows] >= 0)[0]
            if len(sel):
                out.groups.append((pos[grows[sel]], Q[sel].copy(), T[sel].copy()))
        return out

    def put(self, dst_rows, src: "_Approx"):
        """Overwrite rows `dst_rows` of this object with all rows of `src` (best-so-far bookkeeping)."""
        self.sqrt_alpha[dst_rows] = src.sqrt_alpha
        self.logdet[dst_rows] = src.logdet
        self.ok[dst_rows] = src.ok
        self.mu[dst_rows] = src.mu
        # drop the old factors of these rows, then append the new ones
        drop = np.zeros(self.r, dtype=bool)
        drop[dst_rows] = True
        groups = []
        for grows, Q, T in self.groups:
            keep = ~drop[grows]
            if keep.any():
                groups.append((grows[keep], Q[keep], T[keep]))
        for grows, Q, T in src.groups:
            groups.append((np.asarray(dst_rows)[grows], Q, T))
        self.groups = groups


def _empty_approx(r, D):
    out = object.__new__(_Approx)
    out.r, out.D = r, D
    out.sqrt_alpha = np.ones((r, D))
    out.logdet = np.zeros(r)
    out.ok = np.zeros(r, dtype=bool)
    out.mu = np.zeros((r, D))
    out.groups = []
    return out


# ------------------------------------------------------------------------------------------------ PSIS
def _gpdfit(x):
    """Zhang & Stephens (2009) estimate of the generalised Pareto (k, sigma) for sorted exceedances x > 0,
    with the weakly informative prior on k of Vehtari et al. (2024)."""
    n = len(x)
    m_est = 30 + int(n ** 0.5)
    b = 1 - np.sqrt(m_est / (np.arange(1, m_est + 1, dtype=float) - 0.5))
    b /= 3 * x[int(n / 4 + 0.5) - 1]
    b += 1 / x[-1]
    k = np.log1p(-b[:, None] * x).mean(axis=1)
    L = n * (np.log(-(b / k)) - k - 1)
    with np.errstate(over="ignore"):
        w = 1 / np.exp(L - L[:, None]).sum(axis=1)
    keep = w >= 10 * EPS
    w, b = w[keep], b[keep]
    w /= w.sum()
    b_post = (b * w).sum()
    k_post = np.log1p(-b_post * x).mean()
    sigma = -k_post / b_post
    k_post = (n * k_post + 10 * 0.5) / (n + 10)
    return k_post, sigma


def psis_log_weights(log_ratios):
    """Pareto-smoothed, normalised log importance weights and the Pareto k of the tail (Vehtari et al.)."""
    x = np.array(log_ratios, dtype=np.float64)
    n = len(x)
    x[~np.isfinite(x)] = -np.inf
    if not np.isfinite(x.max()):
        return np.full(n, -np.log(n)), np.inf
    x -= x.max()
    k = np.inf
    tail_len = int(np.ceil(min(0.2 * n, 3 * np.sqrt(n))))
    if n > 5 and tail_len > 4:
        order = np.argsort(x)
        cutoff = max(x[order[-tail_len - 1]], np.log(np.finfo(float).tiny))
        tail = np.nonzero(x > cutoff)[0]
        if len(tail) > 4:
            tord = np.argsort(x[tail])
            exc = np.exp(x[tail][tord]) - np.exp(cutoff)
            with np.errstate(all="ignore"):
                k, sigma = _gpdfit(exc)
            if np.isfinite(k) and np.isfinite(sigma) and sigma > 0:
                p = (np.arange(len(tail)) + 0.5) / len(tail)
                q = -np.log1p(-p) if abs(k) < 1e-12 else np.expm1(-k * np.log1p(-p)) / k
                sm = np.log(sigma * q + np.exp(cutoff))
                x[tail[tord]] = np.minimum(sm, 0.0)
            else:
                k = np.inf
    m = x.max()
    x -= m + np.log(np.exp(x - m).sum())
    return x, float(k)


# ------------------------------------------------------------------------------------------------ L-BFGS
def _two_loop(g, S, Y, m):
    """-H g for every row: two-loop recursion over the (up to J) newest pairs; rows with m = 0 get -g."""
    R, J, D = S.shape
    q = g.copy()
    sy = np.einsum("rjd,rjd->rj", S, Y)
    valid = np.arange(J)[None, :] < m[:, None]
    with np.errstate(all="ignore"):
        rho = np.where(valid & (sy > 0), 1.0 / sy, 0.0)
    a = np.zeros((R, J))
    for i in range(J - 1, -1, -1):
        a[:, i] = rho[:, i] * np.einsum("rd,rd->r", S[:, i], q)
        q -= a[:, i, None] * Y[:, i]
    last = np.maximum(m - 1, 0)
    rows = np.arange(R)
    yy = np.einsum("rd,rd->r", Y[rows, last], Y[rows, last])
    with np.errstate(all="ignore"):
        gamma = np.where((m > 0) & (yy > 0), sy[rows, last] / yy, 1.0)
    q *= gamma[:, None]
    for i in range(J):
        bcoef = rho[:, i] * np.einsum("rd,rd->r", Y[:, i], q)
        q += (a[:, i] - bcoef)[:, None] * S[:, i]
    return -q


def _push(S, Y, m, rows, s, y):
    """Append the pair (s, y) to the history of `rows` (oldest pair dropped when full)."""
    J = S.shape[1]
    full = m[rows] >= J
    fr = rows[full]
    if len(fr):
        S[fr, :-1] = S[fr, 1:]
        Y[fr, :-1] = Y[fr, 1:]
        S[fr, -1] = s[full]
        Y[fr, -1] = y[full]
    nf = rows[~full]
    if len(nf):
        S[nf, m[nf]] = s[~full]
        Y[nf, m[nf]] = y[~full]
        m[nf] += 1


def pathfinder(target, *, num_paths: int = 4, num_draws: int = 1000, num_out: int | None = None, seed: int = 0,
               history_size: int = 5, num_elbo_draws: int = 25, max_lbfgs_iters: int = 1000, init_alpha: float = 1e-3,
               tol_obj: float = 1e-12, tol_rel_obj: float = 1e4, tol_grad: float = 1e-8, tol_rel_grad: float = 1e7,
               tol_param: float = 1e-8, init_radius: float = 2.0, psis_resample: bool = True, init=None,
               eval_rows: int | None = None) -> PathfinderResult:
    """Multi-path Pathfinder of every gene of `target` (G genes x `num_paths` paths in one batched L-BFGS).

    `num_draws` draws are taken from the best approximation of every path; `num_out` (default `num_draws`) draws per
    gene are returned after PSIS resampling over the gene's paths (`psis_resample=False` or a single path: the draws
    of each path in turn, unweighted).  `init`: optional (G, P, D) starting points.  `eval_rows` bounds the rows of one target call
    (default: 512 MB of parameter vectors).
    """
    D, G, P, J = target.dim, target.G, int(num_paths), int(history_size)
    eval_rows = max(1, min(1 << 14, (1 << 29) // (8 * D))) if eval_rows is None else int(eval_rows)
    num_out = num_draws if num_out is None else int(num_out)
    R = G * P
    gene_of = np.repeat(np.arange(G, dtype=np.int32), P)
    path_of = np.tile(np.arange(P), G)
    n_lp_grad = n_lp = 0

    def f_grad(theta, go):
        nonlocal n_lp_grad
        f = np.empty(len(theta))
        g = np.empty_like(theta)
        for a in range(0, len(theta), eval_rows):
            lp, gr = target.lp_grad(theta[a:a + eval_rows], go[a:a + eval_rows])
            f[a:a + eval_rows], g[a:a + eval_rows] = -lp, -gr
        n_lp_grad += len(theta)
        bad = ~(np.isfinite(f) & np.isfinite(g).all(axis=1))
        f[bad] = np.inf
        return f, g

    def lp_only(theta, go):
        nonlocal n_lp
        out = np.empty(len(theta))
        for a in range(0, len(theta), eval_rows):
            out[a:a + eval_rows] = target.lp(theta[a:a + eval_rows], go[a:a + eval_rows])
        n_lp += len(theta)
        out[~np.isfinite(out)] = -np.inf
        return out

    # ---- starting points: uniform(-radius, radius), retried until lp and its gradient are finite (Stan: <= 100)
    x = np.empty((R, D))
    rngs = [_row_rng(seed, target.gene_id0 + int(gene_of[r]), int(path_of[r])) for r in range(R)]
    if init is not None:
        x[:] = np.asarray(init, dtype=np.float64).reshape(R, D)
    else:
        for r in range(R):
            x[r] = rngs[r].uniform(-init_radius, init_radius, D)
    f, g = f_grad(x, gene_of)
    for _ in range(100):
        bad = np.nonzero(~np.isfinite(f))[0]
        if not len(bad) or init is not None:
            break
        for r in bad:
            x[r] = rngs[r].uniform(-init_radius, init_radius, D)
        f[bad], g[bad] = f_grad(x[bad], gene_of[bad])
    active = np.isfinite(f)

    S = np.zeros((R, J, D))
    Y = np.zeros((R, J, D))
    m = np.zeros(R, dtype=np.int64)
    alpha = np.ones((R, D))
    best = _empty_approx(R, D)
    best_elbo = np.full(R, -np.inf)
    best_iter = np.zeros(R, dtype=np.int64)
    iters = np.zeros(R, dtype=np.int64)
    ret_code = np.zeros(R, dtype=np.int64)    # 0 running / max iters, 1.. Stan's termination reasons, -1 failed

    for it in range(1, max_lbfgs_iters + 1):
        act = np.nonzero(active)[0]
        if not len(act):
            break
        # ---- search direction
        d = _two_loop(g[act], S[act], Y[act], m[act])
        gd = np.einsum("rd,rd->r", g[act], d)
        notdesc = ~(gd < 0)
        if notdesc.any():
            rr = act[notdesc]
            m[rr] = 0
            S[rr] = 0
            Y[rr] = 0
            d[notdesc] = -g[rr]
            gd[notdesc] = -np.einsum("rd,rd->r", g[rr], g[rr])
        # ---- weak-Wolfe line search by bisection / expansion, all rows in lock step
        n = len(act)
        t = np.where(m[act] > 0, 1.0, init_alpha)
        lo, hi = np.zeros(n), np.full(n, np.inf)
        xn, fn, gn = x[act].copy(), f[act].copy(), g[act].copy()
        have = np.zeros(n, dtype=bool)          # an Armijo point is stored in (xn, fn, gn)
        und = np.ones(n, dtype=bool)
        for _ls in range(40):
            u = np.nonzero(und)[0]
            if not len(u):
                break
            xt = x[act[u]] + t[u, None] * d[u]
            ft, gt = f_grad(xt, gene_of[act[u]])
            armijo = ft <= f[act[u]] + 1e-4 * t[u] * gd[u]
            curv = np.einsum("rd,rd->r", gt, d[u]) >= 0.9 * gd[u]
            store = armijo & (~have[u] | (ft < fn[u]))
            su = u[store]
            xn[su], fn[su], gn[su] = xt[store], ft[store], gt[store]
            have[su] = True
            done = armijo & curv
            hi[u[~armijo]] = t[u[~armijo]]
            lo[u[armijo & ~curv]] = t[u[armijo & ~curv]]
            und[u[done]] = False
            # the best Armijo point must be the accepted one for rows that finish now
            fin = u[done]
            xn[fin], fn[fin], gn[fin] = xt[done], ft[done], gt[done]
            t = np.where(np.isfinite(hi), 0.5 * (lo + hi), 2.0 * np.maximum(lo, t))
            with np.errstate(invalid="ignore"):
                und &= ~np.isfinite(hi) | ((hi - lo) > 1e-16 * np.maximum(hi, 1e-300))   # bracket collapsed
        moved = have | ~und
        moved &= np.isfinite(fn)
        failed = act[~moved]
        active[failed] = False
        ret_code[failed] = -1
        ok = np.nonzero(moved)[0]
        rows = act[ok]
        if not len(rows):
            continue
        s = xn[ok] - x[rows]
        y = gn[ok] - g[rows]
        f_old = f[rows].copy()
        x[rows], f[rows], g[rows] = xn[ok], fn[ok], gn[ok]
        iters[rows] = it
        # ---- curvature check (Stan's check_curve) -> history and diagonal update
        Dk = np.einsum("rd,rd->r", y, s)
        with np.errstate(all="ignore"):
            thetak = np.abs(np.einsum("rd,rd->r", y, y) / Dk)
        acc = (Dk > 0) & (thetak <= 1e12)
        if acc.any():
            ar = rows[acc]
            sa, ya, al = s[acc], y[acc], alpha[ar]
            a_ = np.einsum("rd,rd->r", ya * al, ya)
            b_ = Dk[acc]
            c_ = np.einsum("rd,rd->r", sa / al, sa)
            with np.errstate(all="ignore"):
                new = 1.0 / (a_[:, None] / (b_[:, None] * al) + ya * ya / b_[:, None]
                             - a_[:, None] * sa * sa / (b_[:, None] * c_[:, None] * al * al))
            good = np.isfinite(new).all(axis=1) & (new > 0).all(axis=1)
            alpha[ar[good]] = new[good]
            _push(S, Y, m, ar, sa, ya)
        # ---- normal approximation at the new iterate and its ELBO
        ap = _Approx(x[rows], g[rows], alpha[rows], S[rows], Y[rows], m[rows])
        u_e = _normals(seed, 1, 0, it, (num_elbo_draws, D))
        elbo = np.full(len(rows), -np.inf)
        step = max(1, eval_rows // max(1, num_elbo_draws))
        for a0 in range(0, len(rows), step):
            sub = np.arange(a0, min(a0 + step, len(rows)))
            sub = sub[ap.ok[sub]]
            if not len(sub):
                continue
            part = ap.take(sub)
            phi, logq = part.draw(u_e)
            lpv = lp_only(phi.reshape(-1, D), np.repeat(gene_of[rows[sub]], num_elbo_draws)).reshape(len(sub), -1)
            elbo[sub] = (lpv - logq).mean(axis=1)
        elbo[~np.isfinite(elbo)] = -np.inf
        better = elbo > best_elbo[rows]
        if better.any():
            br = np.nonzero(better)[0]
            best.put(rows[br], ap.take(br))
            best_elbo[rows[br]] = elbo[br]
            best_iter[rows[br]] = it
        # ---- Stan's L-BFGS termination tests
        df = np.abs(f[rows] - f_old)
        gnorm = np.sqrt(np.einsum("rd,rd->r", g[rows], g[rows]))
        snorm = np.sqrt(np.einsum("rd,rd->r", s, s))
        Hg = _two_loop(g[rows], S[rows], Y[rows], m[rows])
        gHg = -np.einsum("rd,rd->r", g[rows], Hg)
        code = np.zeros(len(rows), dtype=np.int64)
        code = np.where(df < tol_obj, 1, code)
        code = np.where((code == 0) & (df / np.maximum(np.maximum(np.abs(f[rows]), np.abs(f_old)), 1.0)
                                       < tol_rel_obj * EPS), 2, code)
        code = np.where((code == 0) & (gnorm < tol_grad), 3, code)
        code = np.where((code == 0) & (gHg / np.maximum(np.abs(f[rows]), 1.0) < tol_rel_grad * EPS), 4, code)
        code = np.where((code == 0) & (snorm < tol_param), 5, code)
        stop = code != 0
        ret_code[rows[stop]] = code[stop]
        active[rows[stop]] = False

    # ---- draws from the best approximation of every path, lp of each, PSIS over the paths of a gene
    out_draws = np.zeros((G, num_out, D))
    out_lp = np.full((G, num_out), -np.inf)
    out_lpa = np.full((G, num_out), -np.inf)
    out_path = np.zeros((G, num_out), dtype=np.int64)
    pareto_k = np.full(G, np.inf)
    blk = max(1, min(num_draws, eval_rows))
    for gi in range(G):
        rows = np.arange(gi * P, (gi + 1) * P)
        okp = [int(r) for r in rows if best.ok[r] and np.isfinite(best_elbo[r])]
        if not okp:
            continue                                       # every path of this gene failed: draws stay zero, lp -inf
        lps, lpas, where = [], [], []
        for r in okp:
            part = best.take(np.array([r]))
            for b0 in range(0, num_draws, blk):
                nb = min(blk, num_draws - b0)
                u = _normals(seed, 2, int(path_of[r]), b0 // blk, (blk, D), target.gene_id0 + gi)[:nb]
                phi, logq = part.draw(u)
                lps.append(lp_only(phi[0], np.full(nb, gi, dtype=np.int32)))
                lpas.append(logq[0])
                where.append(np.stack([np.full(nb, r), np.arange(b0, b0 + nb)], axis=1))
        lps, lpas, where = np.concatenate(lps), np.concatenate(lpas), np.concatenate(where)
        if psis_resample and P > 1:              # single-path Pathfinder returns its draws as they are (Stan)
            logw, pareto_k[gi] = psis_log_weights(lps - lpas)
            rng = _row_rng(seed, target.gene_id0 + gi, 255)
            pick = rng.choice(len(lps), size=num_out, replace=True, p=np.exp(logw))
        else:
            order = np.argsort(where[:, 1], kind="stable")   # draw 0 of every path, then draw 1, ...
            pick = order[np.arange(num_out) % len(order)]
        out_lp[gi], out_lpa[gi], out_path[gi] = lps[pick], lpas[pick], path_of[where[pick, 0]] + 1
        # regenerate only the selected draws
        for r in okp:
            sel = np.nonzero(where[pick, 0] == r)[0]
            if not len(sel):
                continue
            part = best.take(np.array([r]))
            didx = where[pick[sel], 1]
            for b in np.unique(didx // blk):
                u = _normals(seed, 2, int(path_of[r]), int(b), (blk, D), target.gene_id0 + gi)
                inb = didx // blk == b
                phi, _ = part.draw(u[didx[inb] % blk])
                out_draws[gi, sel[inb]] = phi[0]
    return PathfinderResult(out_draws, out_lp, out_lpa, out_path, pareto_k, best_elbo.reshape(G, P),
                            best_iter.reshape(G, P), iters.reshape(G, P), n_lp_grad, n_lp,
                            info={"return_code": ret_code.reshape(G, P)})
