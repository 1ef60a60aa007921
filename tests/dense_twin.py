"""Dense numpy/scipy statement of the model's MATH, written independently of oracle/spamtree_oracle.cpp.
It pins the oracle (the reference has no tests of its own) and gives size-independent checks of the CUDA path."""
import numpy as np


def theta_unpack(theta, q):
    """covariance_functions.cpp:34-92"""
    n_cbase = 3 if q > 2 else 1
    ai1, ai2, phi = theta[:q], theta[q:2 * q], theta[2 * q:3 * q]
    tm = theta[3 * q:3 * q + n_cbase]
    D = np.zeros((q, q), dtype=theta.dtype)
    rest = theta[3 * q + n_cbase:]
    ix = 0
    for j in range(q):
        for i in range(j + 1, q):
            D[i, j] = D[j, i] = rest[ix]
            ix += 1
    return ai1, ai2, phi, tm, D


def cov(coordsA, mvA, coordsB, mvB, theta, q, dtype=np.float64):
    """man/CrossCovarianceAG10.Rd:44-52 / covariance_functions.cpp:113-135,213-286 (mv 1-based).
    dtype = np.longdouble evaluates the same formulas in extended precision (ground truth for the deep blocks)."""
    ai1, ai2, phi, tm, D = theta_unpack(np.asarray(theta, dtype=dtype), q)
    coordsA, coordsB = np.asarray(coordsA, dtype=dtype), np.asarray(coordsB, dtype=dtype)
    h = np.sqrt(((coordsA[:, None, :] - coordsB[None, :, :]) ** 2).sum(-1))
    if q == 1:
        return ai1[0] * np.exp(-tm[0] * h)
    a, b = np.asarray(mvA) - 1, np.asarray(mvB) - 1
    v = D[a[:, None], b[None, :]]
    if q > 2:
        psi = (1 + tm[0] * v) ** tm[1]          # psi(v) = (a v + 1)^beta ; C = exp(-c h / sqrt(psi)) / psi
        base = np.exp(-tm[2] * h / np.sqrt(psi)) / psi
    else:
        base = np.exp(-tm[0] * h / np.sqrt(v + 1)) / (v + 1)
    same = v == 0
    out = ai1[a][:, None] * ai1[b][None, :] * base
    out = np.where(same, ai1[a][:, None] ** 2 * base + ai2[a][:, None] ** 2 * np.exp(-phi[a][:, None] * h), out)
    return out


class Twin:
    """dense per-block quantities for a problem dict from tests/common.make_problem"""

    def __init__(self, pb, theta=None):
        d, t = pb["d"], pb["tree"]
        self.q, self.theta = pb["q"], pb["theta"] if theta is None else theta
        self.coords, self.mv, self.y = d["coords"], d["mv_id"], d["y"]
        ip, ii, pp, pi = t["indexing_ptr"], t["indexing_idx"], t["parents_ptr"], t["parents_idx"]
        self.nb = t["n_blocks"]
        self.rows = [ii[ip[u]:ip[u + 1]] for u in range(self.nb)]
        self.parents = [pi[pp[u]:pp[u + 1]] for u in range(self.nb)]
        self.level = t["block_groups"].astype(int)
        self.obs = np.array([np.isfinite(self.y[r]).sum() for r in self.rows])
        levels = sorted(set(self.level[self.obs > 0]))
        rr = t["res_is_ref"]
        self.isref = np.array([self.obs[u] > 0 and rr[levels.index(self.level[u])] == 1 if self.obs[u] > 0 else False for u in range(self.nb)])
        self.levels = levels

    def K(self, ra, rb):
        return cov(self.coords[ra], self.mv[ra], self.coords[rb], self.mv[rb], self.theta, self.q)

    def prow(self, u):
        return np.concatenate([self.rows[p] for p in self.parents[u]]) if len(self.parents[u]) else np.zeros(0, dtype=np.int64)

    def block(self, u):
        """H (m x P), and for reference blocks Ri (m x m lower) / else the m values 1/sqrt(R_ii)"""
        ru, rp = self.rows[u], self.prow(u)
        Kuu = self.K(ru, ru)
        if rp.size:
            Kpp, Kup = self.K(rp, rp), self.K(ru, rp)
            H = np.linalg.solve(Kpp, Kup.T).T
            R = Kuu - H @ Kup.T
        else:
            H, R = np.zeros((ru.size, 0)), Kuu
        if self.isref[u]:
            Ri = np.linalg.inv(np.linalg.cholesky((R + R.T) / 2))
        else:
            Ri = 1 / np.sqrt(np.diag(R))
        return H, Ri

    def precision(self):
        """Q = sum_u (I - H_u)' prec_u (I - H_u) over observed blocks, dense n_all x n_all; and sum log diag Ri"""
        n = self.y.size
        Q = np.zeros((n, n))
        logdet = 0.0
        self.H, self.Ri = {}, {}
        for u in range(self.nb):
            if self.obs[u] == 0:
                continue
            H, Ri = self.block(u)
            self.H[u], self.Ri[u] = H, Ri
            B = np.zeros((self.rows[u].size, n))
            B[np.arange(self.rows[u].size), self.rows[u]] = 1
            rp = self.prow(u)
            if rp.size:
                B[:, rp] -= H
            prec = Ri.T @ Ri if self.isref[u] else np.diag(Ri ** 2)
            Q += B.T @ prec @ B
            logdet += np.log(np.diag(Ri)).sum() if self.isref[u] else np.log(Ri).sum()
        return Q, logdet

    def loglik(self, w):
        Q, logdet = self.precision()
        nobs = sum(self.rows[u].size for u in range(self.nb) if self.obs[u] > 0)
        return logdet - 0.5 * nobs * np.log(2 * np.pi) - 0.5 * w @ Q @ w

    def gibbs_sweep(self, w, z, tausq_inv_row, resid):
        """one sweep with the reference's message timing: the message of block c to ancestor a is
        formed when c is sampled, from the then-current w.  Returns new w and the per-block (Sigi_tot, Smu_tot)."""
        w = w.copy()
        if not hasattr(self, "H"):
            self.precision()
        msgS = {u: np.zeros((self.rows[u].size,) * 2) for u in range(self.nb)}
        msgM = {u: np.zeros(self.rows[u].size) for u in range(self.nb)}
        probes = {}
        for lev in reversed(self.levels):
            for u in [u for u in range(self.nb) if self.obs[u] > 0 and self.level[u] == lev]:
                ru, rp = self.rows[u], self.prow(u)
                H, Ri = self.H[u], self.Ri[u]
                prec = Ri.T @ Ri if self.isref[u] else np.diag(Ri ** 2)
                tau = tausq_inv_row[ru]
                if self.isref[u]:
                    Sig = prec + msgS[u] + np.diag(tau)
                    Smu = prec @ (H @ w[rp]) + msgM[u] + tau * resid[ru] if rp.size else msgM[u] + tau * resid[ru]
                    Sc = np.linalg.inv(np.linalg.cholesky((Sig + Sig.T) / 2))
                    w[ru] = Sc.T @ (Sc @ Smu + z[ru])
                    probes[u] = (Sig, Smu)
                else:
                    sig = np.diag(prec) + tau
                    smu = np.diag(prec) * (H @ w[rp]) + tau * resid[ru]
                    w[ru] = smu / sig + z[ru] / np.sqrt(sig)
                    probes[u] = (sig, smu)
                off = 0
                for a in self.parents[u]:
                    ma = self.rows[a].size
                    Ha = H[:, off:off + ma]
                    others = np.ones(rp.size, dtype=bool)
                    others[off:off + ma] = False
                    msgS[a] = msgS[a] + Ha.T @ prec @ Ha
                    msgM[a] = msgM[a] + Ha.T @ prec @ (w[ru] - H[:, others] @ w[rp][others])
                    off += ma
        return w, probes


# ---- extended-precision (x87 80-bit long double, eps ~ 1e-19) dense algebra for a block and its ancestor chain
def chol_ld(A):
    A = np.array(A, dtype=np.longdouble)
    n = A.shape[0]
    for j in range(n):
        A[j, j] = np.sqrt(A[j, j] - A[j, :j] @ A[j, :j])
        if j + 1 < n:
            A[j + 1:, j] = (A[j + 1:, j] - A[j + 1:, :j] @ A[j, :j]) / A[j, j]
    return np.tril(A)


def solve_lower_ld(L, B):
    """L^-1 B"""
    X = np.array(B, dtype=np.longdouble)
    for i in range(L.shape[0]):
        X[i] = (X[i] - L[i, :i] @ X[:i]) / L[i, i]
    return X


def solve_upper_ld(U, B):
    """U^-1 B"""
    X = np.array(B, dtype=np.longdouble)
    for i in range(U.shape[0] - 1, -1, -1):
        X[i] = (X[i] - U[i, i + 1:] @ X[i + 1:]) / U[i, i]
    return X


def block_truth_ld(coords, mv, ru, rp, theta, q, isref):
    """H = K_{u,pa} K_pa^-1 (m x P) and Ri = chol(K_uu - H K_{pa,u})^-1 (or 1/sqrt(diag) for a non-reference block) in
    extended precision: App. A of the survey, spamtree_model.cpp:885-897, :931-948"""
    ld = np.longdouble
    Kpp = cov(coords[rp], mv[rp], coords[rp], mv[rp], theta, q, ld)
    Kpu = cov(coords[rp], mv[rp], coords[ru], mv[ru], theta, q, ld)
    Kuu = cov(coords[ru], mv[ru], coords[ru], mv[ru], theta, q, ld)
    L = chol_ld(Kpp)
    Y = solve_lower_ld(L, Kpu)
    H = solve_upper_ld(L.T, Y).T
    R = Kuu - Y.T @ Y
    if isref:
        Ri = solve_lower_ld(chol_ld(R), np.eye(ru.size, dtype=ld))
    else:
        Ri = 1 / np.sqrt(np.diag(R))
    return H, Ri
