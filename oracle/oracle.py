"""ctypes wrapper of the CPU oracle (oracle/spamtree_oracle.cpp).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
arm; never by the product package.  The reference ships no tests of its own; the oracle is pinned against the
reference's own sources compiled here (oracle/_ref, tests/test_oracle_vs_reference.py, tests/test_reference_driver.py): see
the .cpp header.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libspamtree_oracle.so")


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "spamtree_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.or_create.restype = C.c_void_p
        L.or_create.argtypes = [C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp, _ip, C.c_int, _ip, _ip, _ip, _ip, _ip, _ip, _dp,
                                _dp, _ip, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_double, C.c_int]
        L.or_destroy.argtypes = [C.c_void_p]
        L.or_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.or_set_threads.argtypes = [C.c_int]
        L.or_use_blas.restype = C.c_int
        L.or_use_blas.argtypes = [C.c_char_p]
        L.or_max_threads.restype = C.c_int
        L.or_theta_update.argtypes = [C.c_void_p, C.c_int, _dp]
        L.or_build.restype = C.c_int
        L.or_build.argtypes = [C.c_void_p, C.c_int, _dp]
        L.or_loglik_w.argtypes = [C.c_void_p, C.c_int, _dp]
        L.or_gibbs.restype = C.c_int
        L.or_gibbs.argtypes = [C.c_void_p, _dp]
        L.or_swap.argtypes = [C.c_void_p]
        L.or_predict.argtypes = [C.c_void_p, C.c_int]
        L.or_sample_beta.argtypes = [C.c_void_p, _dp]
        L.or_sample_tausq.argtypes = [C.c_void_p, _dp]
        L.or_get_w.argtypes = [C.c_void_p, _dp]
        L.or_set_w.argtypes = [C.c_void_p, _dp]
        L.or_get_params.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.or_set_tausq_inv.argtypes = [C.c_void_p, _dp]
        L.or_get.restype = C.c_int64
        L.or_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, _dp, C.c_int64]
        L.or_cross_covariance_ag10.argtypes = [_dp, _ip, C.c_int64, _dp, _ip, C.c_int64, _dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int, _dp]
        L.or_kthresholds.argtypes = [_dp, C.c_int64, C.c_int, _dp]
        L.or_part_axis_parallel_lmt.argtypes = [_dp, C.c_int64, C.c_int, _dp, _ip, _dp]
        L.or_number_revalue.argtypes = [_ip, C.c_int64, C.c_int, _ip, _ip, C.c_int64, _ip]
        L.or_make_edges.argtypes = [_dp, C.c_int64, C.c_int, _ip, C.c_int64, _ip, C.c_int, _ip, _ip, _ip, _ip, _ip]
        L.or_mcmc.restype = C.c_int
        L.or_mcmc.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_uint64, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.or_timed_iteration.restype = C.c_double
        L.or_timed_iteration.argtypes = [C.c_void_p, _dp, C.c_int]
        _lib = L
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _pd(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _pi(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _cm(a):
    a = np.asarray(a, dtype=np.float64)
    return _f(a.T).reshape(-1) if a.ndim == 2 else _f(a).reshape(-1)


def use_blas(on=True):
    """route the oracle's dense kernels through scipy's bundled OpenBLAS (the "port+blas" CPU baseline of bench.py); returns
    the library path, or None when it is not available.  use_blas(False) restores the plain loops (the parity default)."""
    if not on:
        lib().or_use_blas(None)
        return None
    try:
        import glob
        import scipy
        cands = sorted(glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so")))
    except Exception:
        cands = []
    for c in cands:
        if lib().or_use_blas(os.path.abspath(c).encode()) == 0:
            return os.path.abspath(c)
    return None


FLAG_LEAN, FLAG_Q1_NORM_EXPANSION, FLAG_CORRECT_BETA_INDEX, FLAG_PROBES, FLAG_CORRECT_PREDICT_CACHE = 1, 2, 4, 8, 16


class OracleModel:
    """Restatement of SpamTreeMV (src/spamtree_model.h:22-212).  CSR inputs as in include/spamtree_b200.h."""

    def __init__(self, y, X, coords, mv_id, res_is_ref, csr, limited_tree, block_names, block_groups, beta, theta, tausq,
                 flags=FLAG_PROBES):
        L = lib()
        self.y = _f(y).reshape(-1)
        self.n_all = self.y.size
        Xa = np.asarray(X, dtype=np.float64).reshape(self.n_all, -1)
        self.p = Xa.shape[1]
        self.mv = _i(mv_id)
        self.q = int(np.unique(self.mv).size)
        self._keep = [_cm(Xa), _cm(coords)] + [_i(a) for a in csr] + [_f(block_names), _f(block_groups), _i(res_is_ref), _f(theta), _f(beta)]
        Xc, cc, ip, ii, pp, pi, cp, ci, bn, bg, rr, th, be = self._keep
        self.n_blocks = ip.size - 1
        self.npar = th.size
        self.csr = (ip, ii, pp, pi, cp, ci)
        self.h = L.or_create(self.n_all, self.p, self.q, _pd(self.y), _pd(Xc), _pd(cc), _pi(self.mv), self.n_blocks, _pi(ip), _pi(ii),
                             _pi(pp), _pi(pi), _pi(cp), _pi(ci), _pd(bn), _pd(bg), _pi(rr), rr.size, int(bool(limited_tree)),
                             _pd(th), th.size, _pd(be), float(tausq), int(flags))
        if not self.h:
            raise RuntimeError("oracle: model construction failed")

    def close(self):
        if getattr(self, "h", None):
            lib().or_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def seed(self, s):
        lib().or_seed(self.h, int(s))

    def theta_update(self, slot, theta):
        t = _f(theta)
        lib().or_theta_update(self.h, slot, _pd(t))

    def get_loglik_comps_w(self, slot):
        o = np.zeros(3)
        ok = lib().or_build(self.h, slot, _pd(o))
        return bool(ok), float(o[0]), float(o[1])

    def get_loglik_w(self, slot=0):
        o = np.zeros(2)
        lib().or_loglik_w(self.h, slot, _pd(o))
        return float(o[0]), float(o[1])

    def deal_with_w(self, z=None):
        zz = None if z is None else _f(z)
        if not lib().or_gibbs(self.h, _pd(zz)):
            raise RuntimeError("Error at gibbs_sample_w")

    def accept_make_change(self):
        lib().or_swap(self.h)

    def predict(self, theta_changed=True):
        lib().or_predict(self.h, int(bool(theta_changed)))

    def gibbs_sample_beta(self, zb=None):
        z = None if zb is None else _cm(zb)
        lib().or_sample_beta(self.h, _pd(z))

    def gibbs_sample_tausq(self, fixed=None):
        f = None if fixed is None else _f(fixed)
        lib().or_sample_tausq(self.h, _pd(f))

    @property
    def w(self):
        o = np.zeros(self.n_all)
        lib().or_get_w(self.h, _pd(o))
        return o

    @w.setter
    def w(self, v):
        a = _f(v)
        lib().or_set_w(self.h, _pd(a))

    def params(self):
        B, t, xb = np.zeros(self.p * self.q), np.zeros(self.q), np.zeros(self.n_all)
        lib().or_get_params(self.h, _pd(B), _pd(t), _pd(xb))
        return {"Bcoeff": B.reshape(self.q, self.p).T.copy(), "tausq_inv": t, "XB": xb}

    def set_tausq_inv(self, t):
        a = _f(t)
        lib().or_set_tausq_inv(self.h, _pd(a))

    def get(self, name, u=0, slot=0, c=0):
        n = lib().or_get(self.h, name.encode(), slot, int(u), int(c), None, 0)
        if n < 0:
            raise KeyError(name)
        o = np.zeros(max(n, 1))
        lib().or_get(self.h, name.encode(), slot, int(u), int(c), _pd(o), o.size)
        return o[:n]

    def geti(self, name, u=0, slot=0, c=0):
        return self.get(name, u, slot, c).astype(np.int64)

    def mat(self, name, u, rows, slot=0):
        """column-major matrix probe reshaped to (rows, -1)"""
        v = self.get(name, u, slot)
        return v.reshape(-1, rows).T.copy() if v.size else v.reshape(rows, 0)

    def mcmc(self, bounds, mcmcsd, keep, burn, thin, adapting=True, sample_beta=True, sample_tausq=True, sample_theta=True,
             sample_w=True, sample_predicts=True, seed=1, save_rows=True):
        b, sd = _cm(bounds), _cm(mcmcsd)
        beta, tausq, theta = np.zeros(self.p * keep * self.q), np.zeros(self.q * keep), np.zeros(self.npar * keep)
        w = np.zeros(self.n_all * keep) if save_rows else None
        yh = np.zeros(self.n_all * keep) if save_rows else None
        psd, stats = np.zeros(self.npar ** 2), np.zeros(3)
        rc = lib().or_mcmc(self.h, _pd(b), _pd(sd), keep, burn, thin, int(adapting), int(sample_beta), int(sample_tausq),
                           int(sample_theta), int(sample_w), int(sample_predicts), int(seed), _pd(beta), _pd(tausq), _pd(theta),
                           _pd(w), _pd(yh), _pd(psd), _pd(stats))
        if rc:
            raise RuntimeError(f"oracle mcmc failed ({rc})")
        return {
            "w_mcmc": None if w is None else w.reshape(keep, self.n_all).T.copy(),
            "yhat_mcmc": None if yh is None else yh.reshape(keep, self.n_all).T.copy(),
            "beta_mcmc": beta.reshape(self.q, keep, self.p).transpose(2, 1, 0).copy(),
            "tausq_mcmc": tausq.reshape(keep, self.q).T.copy(),
            "theta_mcmc": theta.reshape(keep, self.npar).T.copy(),
            "paramsd": psd.reshape(self.npar, self.npar).T.copy(),
            "n_accepted": int(stats[0]), "n_chol_fail": int(stats[1]), "mcmc_time": float(stats[2]),
        }

    def timed_iteration(self, theta_prop, do_swap=False):
        t = _f(theta_prop)
        return lib().or_timed_iteration(self.h, _pd(t), int(bool(do_swap)))


def cross_covariance_ag10(coords1, mv1, coords2, mv2, ai1, ai2, phi_i, thetamv, Dmat):
    c1, c2 = np.asarray(coords1, dtype=np.float64), np.asarray(coords2, dtype=np.float64)
    n1, n2 = c1.shape[0], c2.shape[0]
    Dm = np.asarray(Dmat, dtype=np.float64)
    q = Dm.shape[1]
    a, b, m1, m2 = _cm(c1), _cm(c2), _i(mv1), _i(mv2)
    A1, A2, PH, TM, DD = _f(ai1), _f(ai2), _f(phi_i), _f(np.atleast_1d(thetamv)), _cm(Dm)
    out = np.zeros(n1 * n2)
    lib().or_cross_covariance_ag10(_pd(a), _pi(m1), n1, _pd(b), _pi(m2), n2, _pd(A1), _pd(A2), _pd(PH), _pd(TM), TM.size, _pd(DD), q, _pd(out))
    return out.reshape(n2, n1).T.copy()


def kthresholds(x, k):
    x = _f(x)
    res = np.zeros(max(k - 1, 0))
    lib().or_kthresholds(_pd(x), x.size, k, _pd(res))
    return res


def part_axis_parallel_lmt(coords, thresholds):
    coords = np.asarray(coords, dtype=np.float64)
    n, d = coords.shape
    ptr = np.zeros(d + 1, dtype=np.int64)
    for j in range(d):
        ptr[j + 1] = ptr[j] + len(thresholds[j])
    thr = _f(np.concatenate([np.asarray(t, dtype=np.float64) for t in thresholds]) if ptr[-1] else np.zeros(1))
    cm, out = _cm(coords), np.zeros(n * d)
    lib().or_part_axis_parallel_lmt(_pd(cm), n, d, _pd(thr), _pi(ptr), _pd(out))
    return out.reshape(d, n).T.copy()


def number_revalue(original_mat, from_val, to_val):
    om = np.asarray(original_mat, dtype=np.int64)
    nr, nc = om.shape
    flat, fv, tv = _i(om.T.reshape(-1)), _i(from_val), _i(to_val)
    out = np.zeros(nr * nc, dtype=np.int64)
    lib().or_number_revalue(_pi(flat), nr, nc, _pi(fv), _pi(tv), fv.size, _pi(out))
    return out.reshape(nc, nr).T.copy()


def make_edges(parchimat, non_empty_blocks, res_is_ref, limited=False):
    pm = np.asarray(parchimat, dtype=np.float64)
    nr, L = pm.shape
    flat, ne, rr = _cm(pm), _i(non_empty_blocks), _i(res_is_ref)
    counts = np.zeros(3, dtype=np.int64)
    lib().or_make_edges(_pd(flat), nr, L, _pi(ne), ne.size, _pi(rr), int(limited), None, None, None, None, _pi(counts))
    nb = int(counts[0])
    pp, pi = np.zeros(nb + 1, np.int64), np.zeros(max(int(counts[1]), 1), np.int64)
    cp, ci = np.zeros(nb + 1, np.int64), np.zeros(max(int(counts[2]), 1), np.int64)
    lib().or_make_edges(_pd(flat), nr, L, _pi(ne), ne.size, _pi(rr), int(limited), _pi(pp), _pi(pi), _pi(cp), _pi(ci), _pi(counts))
    return {"parents": [pi[pp[i]:pp[i + 1]].copy() for i in range(nb)], "children": [ci[cp[i]:cp[i + 1]].copy() for i in range(nb)],
            "ptrs": (pp, pi[:counts[1]], cp, ci[:counts[2]])}
