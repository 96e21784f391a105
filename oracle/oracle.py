"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy, IEEE float64) of the
reference's acquisition hot path.  Imported only by `tests/`, by
`__graft_entry__.smoke()` and by `bench.py`'s cpu_baseline / `--impl reference`
legs as the *checker / baseline*.  The product package `optimobo_b200` never
imports anything from here and has no CPU fallback.

Pinning status (also in DESIGN.md):
  * acquisition arithmetic (rows a2-a11 of SURVEY.md section 8a): PINNED -- every
    function below is checked against the reference's own code executed in the
    build container (oracle/ref_loader.py, oracle/make_golden.py) and pinned by
    tests/test_oracle_golden.py against the committed fixtures minted from that
    code (tests/golden/acq_golden.npz) and against the hand anchors of SURVEY 8c.
  * GP posterior (row a1): PARITY UNPINNED by the reference -- the arithmetic
    lives in GPy (gpy>=1.10.0, requirements.txt:5, un-vendored, not
    installable offline).  `gp_fit_state`/`gp_posterior` restate GPy's
    published algorithm (kern/src/stationary.py, exact_gaussian_inference.py,
    posterior.py) and are cross-checked against sklearn's
    GaussianProcessRegressor, which *is* installed (tests/test_oracle_gp.py on
    the CPU; tests/test_gpu_pins.py compares the CUDA posterior with sklearn
    DIRECTLY -- the surface util_functions.py:265 calls).

All citations are file:line into /root/reference/optimobo unless noted.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg as sla
from scipy.special import ndtr

SQRT5 = np.sqrt(5.0)
_NORM_PDF_C = np.sqrt(2.0 * np.pi)

KERNEL_MATERN52 = 0
KERNEL_RBF = 1


# =============================================================================
# a1 / a14 -- GP posterior (GPy semantics)
# =============================================================================
def scaled_dist(A, B, ell, form="gpy"):
    """r_ij = || (a_i - b_j) / ell ||.
    form="gpy": GPy stationary.py::_scaled_dist/_unscaled_dist: inputs divided by
    ell, r^2 = |a|^2 + |b|^2 - 2 a.b clipped at 0.  form="direct": explicit
    differences (sklearn kernels.py:1714 cdist; what the CUDA path does)."""
    A = np.asarray(A, float) / np.asarray(ell, float)
    B = np.asarray(B, float) / np.asarray(ell, float)
    if form == "gpy":
        a2 = np.sum(np.square(A), 1)
        b2 = np.sum(np.square(B), 1)
        r2 = -2.0 * A @ B.T + (a2[:, None] + b2[None, :])
        if A is B or (A.shape == B.shape and np.array_equal(A, B)):
            r2[np.diag_indices(min(r2.shape))] = 0.0
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)
    diff = A[:, None, :] - B[None, :, :]
    return np.sqrt(np.sum(diff * diff, -1))


def k_of_r(r, sigma_f2, kernel=KERNEL_MATERN52):
    """GPy Matern52.K_of_r: variance*(1+sqrt5 r+5/3 r^2) exp(-sqrt5 r);
    RBF.K_of_r: variance*exp(-0.5 r^2)."""
    if kernel == KERNEL_MATERN52:
        return sigma_f2 * (1.0 + SQRT5 * r + 5.0 / 3.0 * r ** 2) * np.exp(-SQRT5 * r)
    return sigma_f2 * np.exp(-0.5 * r ** 2)


def gp_fit_state(X, y, ell, sigma_f2, sigma_n2=0.0, jitter=1e-8,
                 kernel=KERNEL_MATERN52, form="gpy"):
    """GPy exact_gaussian_inference.py: Ky = K + (sigma_n2 + 1e-8) I; LW = chol(Ky);
    alpha = dpotrs(LW, y).  Reference call sites: optimisers.py:226-231 etc.
    (Gaussian_noise.variance.fix(0) => sigma_n2 = 0)."""
    X = np.asarray(X, float)
    y = np.asarray(y, float).reshape(-1)
    ell = np.broadcast_to(np.asarray(ell, float), (X.shape[1],)).copy()
    K = k_of_r(scaled_dist(X, X, ell, form), sigma_f2, kernel)
    Ky = K + (sigma_n2 + jitter) * np.eye(len(X))
    L = sla.cholesky(Ky, lower=True)
    alpha = sla.cho_solve((L, True), y)
    return dict(X=X, y=y, ell=ell, sigma_f2=float(sigma_f2), sigma_n2=float(sigma_n2),
                jitter=float(jitter), kernel=kernel, L=L, alpha=alpha, form=form)


def gp_posterior(state, Xnew, chunk=8192, var_floor=1e-15):
    """GPy posterior.py::_raw_predict + gaussian.py::predictive_values:
    mu = Kx^T alpha; var = clip(Kxx - sum((L^-1 Kx)^2), 1e-15, inf) + sigma_n2.
    Returns (mu (m,), var (m,))."""
    Xnew = np.atleast_2d(np.asarray(Xnew, float))
    m = len(Xnew)
    mu = np.empty(m)
    var = np.empty(m)
    for s in range(0, m, chunk):
        xs = Xnew[s:s + chunk]
        Kx = k_of_r(scaled_dist(state["X"], xs, state["ell"], state["form"]),
                    state["sigma_f2"], state["kernel"])          # (n, mc)
        mu[s:s + chunk] = Kx.T @ state["alpha"]
        tmp = sla.solve_triangular(state["L"], Kx, lower=True, check_finite=False)
        v = state["sigma_f2"] - np.sum(np.square(tmp), 0)
        var[s:s + chunk] = np.clip(v, var_floor, np.inf) + state["sigma_n2"]
    return mu, var


def gp_posterior_joint(state, Xnew):
    """Joint posterior (mu (m,), Sigma (m,m)) -- GPy posterior.py::_raw_predict(full_cov=True):
    Sigma = Kxx - tmp^T tmp with tmp = dtrtrs(LW, Kx); the call site is GP.posterior_samples
    (turbo.py:116)."""
    Xnew = np.atleast_2d(np.asarray(Xnew, float))
    Kx = k_of_r(scaled_dist(state["X"], Xnew, state["ell"], state["form"]), state["sigma_f2"], state["kernel"])
    mu = Kx.T @ state["alpha"]
    tmp = sla.solve_triangular(state["L"], Kx, lower=True, check_finite=False)
    Kxx = k_of_r(scaled_dist(Xnew, Xnew, state["ell"], state["form"]), state["sigma_f2"], state["kernel"])
    np.fill_diagonal(Kxx, state["sigma_f2"])
    return mu, Kxx - tmp.T @ tmp


def posterior_samples(state, Xnew, Z, diag_add):
    """mu + chol(Sigma + diag_add I) Z: the Cholesky form of a joint draw (GPy draws through
    np.random.multivariate_normal; the two agree in distribution, and the Cholesky form is what a
    given Z pins down)."""
    mu, S = gp_posterior_joint(state, Xnew)
    Lc = np.linalg.cholesky(S + diag_add * np.eye(len(mu)))
    return mu[:, None] + Lc @ np.asarray(Z, float)


def gp_posterior_std(state, Xnew):
    """sklearn surface (util_functions.py:265): (mean (m,), std (m,)), negative
    variances zeroed (sklearn _gpr.py:480-500)."""
    mu, var = gp_posterior(state, Xnew, var_floor=-np.inf)
    return mu, np.sqrt(np.maximum(var, 0.0))


# =============================================================================
# candidate pool: counter-based generator (new design).  The CUDA
# generator (csrc/candidates.cuh) must be bit-identical.
# =============================================================================
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def counter_uniform(seed, start, count, d):
    """u[i, j] in [0,1): splitmix64 finaliser of (global_index*d + j) keyed by seed."""
    with np.errstate(over="ignore"):
        idx = (np.arange(start, start + count, dtype=np.uint64)[:, None] * np.uint64(d)
               + np.arange(d, dtype=np.uint64)[None, :])
        z = idx + (np.uint64(seed) + np.uint64(1)) * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def candidates_from_counter(seed, start, count, lo, hi):
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    u = counter_uniform(seed, start, count, len(lo))
    return lo[None, :] + (hi - lo)[None, :] * u


# =============================================================================
# a2 -- change()
# =============================================================================
def change_batched(mu, var0, cache, exact_var=None):
    """util_functions.py:217-237.  mu (m,k), var0 (m,) = variance of MODEL 0 (the
    reference scales every objective by model 0's variance, :233).  cache (S,k).
    exact_var (m,k) switches to each model's own variance ("exact" semantics).
    Returns Y (m,S,k)."""
    mu = np.asarray(mu, float)
    cache = np.asarray(cache, float)
    if exact_var is None:
        sd = np.sqrt(np.asarray(var0, float))[:, None, None]
    else:
        sd = np.sqrt(np.asarray(exact_var, float))[:, None, :]
    return cache[None, :, :] * sd + mu[:, None, :]


# =============================================================================
# a3 / a4 -- EHVI 2-D
# =============================================================================
def _pdf(t):
    return np.exp(-t ** 2 / 2.0) / _NORM_PDF_C      # scipy _norm_pdf


def _psi(a, b, m, s):
    """util_functions.py:130-133."""
    t = (b - m) / s
    return s * _pdf(t) + (a - m) * ndtr(t)


def ehvi_stripes(PF, r):
    """util_functions.py:93-112: sort by f2 ascending, pad with (r0,-inf) and
    (-inf, r1).  Returns y1, y2 of length P+2."""
    PF = np.asarray(PF, float)
    idx = np.argsort(PF[:, 1])
    S = PF[idx]
    y1 = np.concatenate(([r[0]], S[:, 0], [-np.inf]))
    y2 = np.concatenate(([-np.inf], S[:, 1], [r[1]]))
    return y1, y2


def ehvi2d_aux_batched(PF, r, mu0, mu1, s0, s1, exact=False):
    """util_functions.py:81-128, vectorised over candidates.
    reference (exact=False): stripes i = 1..P (:120) -- the (P+1)-th stripe of the
    textbook formula is missing, reproduced faithfully.
    exact=True: adds the missing stripe (limit form: y1[P+1] = -inf)."""
    with np.errstate(all="ignore"):
        y1, y2 = ehvi_stripes(PF, r)
        P = len(y1) - 2
        mu0 = np.asarray(mu0, float); mu1 = np.asarray(mu1, float)
        s0 = np.asarray(s0, float); s1 = np.asarray(s1, float)
        sum1 = np.zeros_like(mu0)
        sum2 = np.zeros_like(mu0)
        for i in range(1, P + 1):
            t = (y1[i] - mu0) / s0
            psi2 = _psi(y2[i], y2[i], mu1, s1)
            sum1 = sum1 + (y1[i - 1] - y1[i]) * ndtr(t) * psi2
            sum2 = sum2 + (_psi(y1[i - 1], y1[i - 1], mu0, s0)
                           - _psi(y1[i - 1], y1[i], mu0, s0)) * psi2
        if exact:
            i = P + 1
            psi2 = _psi(y2[i], y2[i], mu1, s1)
            sum2 = sum2 + _psi(y1[i - 1], y1[i - 1], mu0, s0) * psi2
        return sum1 + sum2


def cache_cov(cache):
    """C = np.cov(cache[:,0], cache[:,1]) (ddof=1) -- the constants behind
    util_functions.py:163."""
    C = np.cov(np.asarray(cache)[:, 0], np.asarray(cache)[:, 1])
    return float(C[0, 0]), float(C[0, 1])


def ehvi_batched(mu0, mu1, var0, var1, PF, r, cache, semantics="reference"):
    """util_functions.py:136-167 (EHVI).  reference: 'sigma' handed to EHVI_2D_aux is
    the flattened sample covariance: s0 = var0*C00, s1 = var0*C01."""
    if semantics == "reference":
        c00, c01 = cache_cov(cache)
        var0 = np.asarray(var0, float)
        return ehvi2d_aux_batched(PF, r, mu0, mu1, var0 * c00, var0 * c01, exact=False)
    return ehvi2d_aux_batched(PF, r, mu0, mu1, np.sqrt(var0), np.sqrt(var1), exact=True)


# =============================================================================
# a5 -- crude 3-D EHVI
# =============================================================================
def ehvi3d_batched(mu, var, r, sminus, cache, semantics="reference"):
    """util_functions.py:170-214: mean_j max(0, prod_k (r_k - Y_jk) - Sminus).
    Samples outside the reference box make pygmo raise in the reference
    (unverifiable offline); they contribute 0 here (documented divergence)."""
    mu = np.asarray(mu, float)
    var = np.asarray(var, float)
    Y = change_batched(mu, var[:, 0], cache, exact_var=None if semantics == "reference" else var)
    r = np.asarray(r, float)
    diff = r[None, None, :] - Y
    inside = np.all(diff >= 0, axis=2)
    hvol = np.prod(diff, axis=2) - sminus
    contrib = np.where(inside & (hvol > 0), hvol, 0.0)
    return contrib.sum(1) / Y.shape[1]


# =============================================================================
# a7 -- the twelve scalarisations, vectorised over the leading axes
# =============================================================================
SCALARISATIONS = ["WeightedSum", "Tchebicheff", "AugmentedTchebicheff", "ModifiedTchebicheff",
                  "ExponentialWeightedCriterion", "WeightedNorm", "WeightedPower",
                  "WeightedProduct", "PBI", "IPBI", "QPBI", "APD"]

DEFAULT_PARAMS = {
    "AugmentedTchebicheff": dict(alpha=0.0001),          # scalarisations.py:88
    "ModifiedTchebicheff": dict(alpha=1),                # :126
    "ExponentialWeightedCriterion": dict(p=100),         # :161
    "WeightedNorm": dict(p=3),                           # :185
    "WeightedPower": dict(p=3),                          # :208
    "PBI": dict(theta=5),                                # :252
    "IPBI": dict(theta=5),                               # :287
    "QPBI": dict(theta=5, alpha=5.0, H=5.0),             # :318
    "APD": dict(FE=1, FE_max=10, gamma=0.010304664101210016),  # :360
}


def scalarise(name, F, w, ideal, maxp, **params):
    """F (..., k) -> (...).  Each branch cites the `_do` it restates."""
    with np.errstate(all="ignore"):
        p = dict(DEFAULT_PARAMS.get(name, {}))
        p.update(params)
        F = np.asarray(F, float)
        w = np.asarray(w, float)
        ideal = np.asarray(ideal, float)
        maxp = np.asarray(maxp, float)
        Fp = (F - ideal) / (maxp - ideal)
        k = F.shape[-1]
        if name == "WeightedSum":                       # :43-50
            return np.sum(Fp * w, -1)
        if name == "Tchebicheff":                       # :66-72
            return np.max(w * Fp, -1)
        if name == "AugmentedTchebicheff":              # :92-109
            a = np.abs(Fp)
            return np.max(a * w, -1) + p["alpha"] * np.sum(a, -1)
        if name == "ModifiedTchebicheff":               # :130-149
            a = np.abs(Fp)
            return np.max((a + p["alpha"] * np.sum(a, -1, keepdims=True)) * w, -1)
        if name == "ExponentialWeightedCriterion":      # :165-173
            return np.sum(np.exp(p["p"] * w - 1) * np.exp(p["p"] * Fp), -1)
        if name == "WeightedNorm":                      # :190-197
            return np.power(np.sum(np.power(np.abs(Fp), p["p"]) * w, -1), 1 / p["p"])
        if name == "WeightedPower":                     # :212-219
            return np.sum((Fp ** p["p"]) * w, -1)
        if name == "WeightedProduct":                   # :230-238
            return np.prod((Fp + 100000) ** w, -1)
        if name in ("PBI", "IPBI", "QPBI"):             # :257-273, :291-310, :325-351
            wn = w / np.linalg.norm(w)
            d1 = np.sum(Fp * wn, -1)
            d2 = np.linalg.norm(Fp - d1[..., None] * wn, axis=-1)
            if name == "PBI":
                return d1 + p["theta"] * d2
            if name == "IPBI":
                return p["theta"] * d2 - d1
            d_star = p["alpha"] * (np.reciprocal(float(p["H"])) * np.reciprocal(float(k))
                                   * np.sum(maxp - ideal))
            return d1 + p["theta"] * d2 * (d2 / d_star)
        if name == "APD":                               # :375-397
            tf = Fp.copy()
            zero_rows = np.all(tf == 0, -1)
            norm_f = np.linalg.norm(tf, axis=-1)        # computed BEFORE the 1e-5 patch (:384)
            tf[zero_rows] = 1e-5
            wv = w if not np.all(w == 0) else np.full(k, 1e-5)
            u1 = tf / np.linalg.norm(tf, axis=-1, keepdims=True)
            u2 = wv / np.linalg.norm(wv)
            theta = np.arccos(np.clip(np.sum(u1 * u2, -1), -1.0, 1.0))
            return (1 + k * (p["FE"] / p["FE_max"]) * theta / p["gamma"]) * norm_f
        raise KeyError(name)


# =============================================================================
# a6 -- expected decomposition
# =============================================================================
def expected_decomposition_batched(mu, var, weights, name, ideal, maxp, g_min, cache,
                                   semantics="reference", **params):
    """util_functions.py:285-327: mean_j max(0, g_min - g(Y_j, w)) (the S x S
    broadcast at :324 leaves the mean unchanged)."""
    var = np.asarray(var, float)
    Y = change_batched(mu, var[:, 0], cache, exact_var=None if semantics == "reference" else var)
    g = scalarise(name, Y, weights, ideal, maxp, **params)       # (m,S)
    with np.errstate(all="ignore"):
        return np.mean(np.maximum(0.0, g_min - g), axis=1)


# =============================================================================
# a8-a10 -- EI family
# =============================================================================
def expected_improvement(mu, var, best, var_eps=0.0):
    """optimisers.py:325-344 (var_eps=0), cparego.py:450-469 (0), parego.py:126-145 and
    keep.py:118-137 (1e-6)."""
    with np.errstate(all="ignore"):
        s = np.sqrt(np.asarray(var, float) + var_eps)
        g = (best - np.asarray(mu, float)) / (s + 1e-10)
        return s * (g * ndtr(g) + _pdf(g))


def probability_of_feasibility(mu, var):
    """cparego.py:471-484."""
    with np.errstate(all="ignore"):
        return ndtr((0.0 - np.asarray(mu, float)) / np.sqrt(np.asarray(var, float) + 1e-5))


def constrained_ei(mu, var, best):
    """cparego.py:486-496: mu/var (m, 1+n_constr), column 0 = aggregate model."""
    ei = expected_improvement(mu[:, 0], var[:, 0], best, 0.0)
    pof = np.prod(probability_of_feasibility(mu[:, 1:], var[:, 1:]), axis=1)
    return ei * pof


def pareto_ei(mu_pareto, mu_scalar, var_scalar, best):
    """keep.py:142-150."""
    return mu_pareto * expected_improvement(mu_scalar, var_scalar, best, 1e-6)


# =============================================================================
# a11 -- hypervolume-based PoI (EMO)
# =============================================================================
def hv_poi_batched(mu, var, cells):
    """emo.py:192-228 / vol5 :176-189.  cells (n_cells,2,k): [c][0] upper, [c][1] lower."""
    with np.errstate(all="ignore"):
        mu = np.asarray(mu, float)[:, :2]
        sd = np.sqrt(np.asarray(var, float)[:, :2] + 1e-5)
        cells = np.asarray(cells, float)
        U = cells[:, 0, :][None]      # (1,c,2)
        Lo = cells[:, 1, :][None]
        m_ = mu[:, None, :]
        s_ = sd[:, None, :]
        # scipy norm.cdf(x, loc, scale) = ndtr((x-loc)/scale)
        p = ndtr((U - m_) / s_) - ndtr((Lo - m_) / s_)
        poi = np.sum(np.prod(p, -1), -1)
        valid = np.all(U > m_, -1)
        vol = np.prod(U - np.maximum(Lo, m_), -1)
        imp = np.sum(np.where(valid, vol, 0.0), -1)
        return poi * imp


# =============================================================================
# a13 -- hoistable host prep the reference gets from pygmo / pymoo
# =============================================================================
def calc_pf(Y):
    """util_functions.py:64-77 (pygmo first front, original order; len<2 -> input)."""
    Y = np.asarray(Y, float)
    if len(Y) < 2:
        return Y
    dom = (np.all(Y[:, None, :] <= Y[None, :, :], -1) & np.any(Y[:, None, :] < Y[None, :, :], -1))
    keep = ~np.any(dom, axis=0)
    return Y[keep]


def hypervolume(points, ref):
    """pygmo hypervolume(points).compute(ref) / pymoo HV(ref_point)(Y): exact, 2-D or 3-D.
    Points outside the reference box are dropped (pymoo behaviour)."""
    P = np.asarray(points, float)
    ref = np.asarray(ref, float)
    P = P[np.all(P <= ref, 1)]
    if len(P) == 0:
        return 0.0
    P = calc_pf(P) if len(P) > 1 else P
    k = P.shape[1]
    if k == 2:
        Q = P[np.argsort(P[:, 0])]
        hv, best = 0.0, ref[1]
        for a, b in Q:
            if b < best:
                hv += (ref[0] - a) * (best - b)
                best = b
        return float(hv)
    order = np.argsort(P[:, -1])
    Q = P[order]
    hv = 0.0
    for i in range(len(Q)):
        z_hi = Q[i + 1, -1] if i + 1 < len(Q) else ref[-1]
        if z_hi > Q[i, -1]:
            hv += hypervolume(Q[: i + 1, :-1], ref[:-1]) * (z_hi - Q[i, -1])
    return float(hv)


def decompose_into_cells_2d(pf, ideal, maxp):
    """emo.py:55-152 / util_functions.py:414-517 for a mutually non-dominated 2-D
    front: the sorted-front staircase.  Cell c has upper (f1_c, f2_{c-1}) and
    lower (f1_{c-1}, ideal_2); first cell upper = (f1_0, max(f2_0, max_2)), lower =
    ideal; the last cell's upper f1 = max(f1_last, max_1)."""
    pf = np.asarray(pf, float)
    S = pf[np.argsort(pf[:, 0], kind="stable")]
    n = len(S)
    cells = np.zeros((n + 1, 2, 2))
    cells[0, 0] = [S[0, 0], max(S[0, 1], maxp[1])]
    cells[0, 1] = ideal
    for c in range(1, n + 1):
        f1_hi = S[c, 0] if c < n else max(S[n - 1, 0], maxp[0])
        cells[c, 0] = [f1_hi, S[c - 1, 1]]
        cells[c, 1] = [S[c - 1, 0], ideal[1]]
    return cells


def das_dennis(n_partitions, n_dim):
    """pymoo get_reference_directions("das-dennis", n_dim, n_partitions=...): simplex lattice."""
    out = []

    def rec(prefix, left, depth):
        if depth == n_dim - 1:
            out.append(prefix + [left])
            return
        for i in range(left + 1):
            rec(prefix + [i], left - i, depth + 1)

    rec([], n_partitions, 0)
    return np.asarray(out, float) / n_partitions


def argmax_lowest_index(v):
    """np.argmax semantics (first maximal index); NaNs are treated as -inf."""
    v = np.where(np.isnan(v), -np.inf, np.asarray(v, float))
    return int(np.argmax(v))
