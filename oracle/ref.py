"""ctypes wrapper of oracle/_ref/libspamtree_ref.so — the reference's OWN model layer (spamtree_model.cpp,
covariance_functions.cpp, tree_utils.cpp, tree_dep.cpp, mh_adapt.cpp), compiled unmodified from /root/reference/src
against the Armadillo/Rcpp stand-in in oracle/refshim/.  TEST INFRASTRUCTURE ONLY: it pins the CPU oracle.
The library is built by `make -C oracle ref` where /root/reference exists and travels prebuilt elsewhere."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libspamtree_ref.so")
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
_lib = None


def available():
    if not os.path.exists(LIB_PATH) and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            return False
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp, _ip, C.c_int, _ip, _ip, _ip, _ip, _ip, _ip, _dp, _dp, _ip,
                                 C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_double]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_theta_update.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_build.restype = C.c_int
        L.ref_build.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_loglik_w.argtypes = [C.c_void_p, C.c_int, _dp]
        L.ref_gibbs.restype = C.c_int
        L.ref_gibbs.argtypes = [C.c_void_p, _dp]
        L.ref_swap.argtypes = [C.c_void_p]
        L.ref_predict.argtypes = [C.c_void_p, C.c_int]
        L.ref_sample_beta.argtypes = [C.c_void_p, _dp]
        L.ref_sample_tausq.argtypes = [C.c_void_p, _dp]
        L.ref_get_w.argtypes = [C.c_void_p, _dp]
        L.ref_set_w.argtypes = [C.c_void_p, _dp]
        L.ref_set_tausq_inv.argtypes = [C.c_void_p, _dp]
        L.ref_get_params.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.ref_get.restype = C.c_int64
        L.ref_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, _dp, C.c_int64]
        L.ref_kthresholds.argtypes = [_dp, C.c_int64, C.c_int, _dp]
        L.ref_cross_covariance_ag10.argtypes = [_dp, _ip, C.c_int64, _dp, _ip, C.c_int64, _dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int, _dp]
        L.ref_number_revalue.argtypes = [_ip, C.c_int64, C.c_int, _ip, _ip, C.c_int64, _ip]
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


class RefModel:
    """the reference's SpamTreeMV (src/spamtree_model.h:22-212), same call surface as oracle.OracleModel"""

    def __init__(self, y, X, coords, mv_id, res_is_ref, csr, limited_tree, block_names, block_groups, beta, theta, tausq):
        L = lib()
        self.y = _f(y).reshape(-1)
        self.n_all = self.y.size
        Xa = np.asarray(X, dtype=np.float64).reshape(self.n_all, -1)
        self.p, self.mv = Xa.shape[1], _i(mv_id)
        self.q = int(np.unique(self.mv).size)
        self._keep = [_cm(Xa), _cm(coords)] + [_i(a) for a in csr] + [_f(block_names), _f(block_groups), _i(res_is_ref), _f(theta), _f(beta)]
        Xc, cc, ip, ii, pp, pi, cp, ci, bn, bg, rr, th, be = self._keep
        self.n_blocks = ip.size - 1
        self.h = L.ref_create(self.n_all, self.p, self.q, _pd(self.y), _pd(Xc), _pd(cc), _pi(self.mv), self.n_blocks, _pi(ip), _pi(ii),
                              _pi(pp), _pi(pi), _pi(cp), _pi(ci), _pd(bn), _pd(bg), _pi(rr), rr.size, int(bool(limited_tree)), _pd(th),
                              th.size, _pd(be), float(tausq))
        if not self.h:
            raise RuntimeError("reference model construction failed")

    def close(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def seed(self, s):
        lib().ref_seed(self.h, int(s))

    def theta_update(self, slot, theta):
        t = _f(theta)
        lib().ref_theta_update(self.h, slot, _pd(t))

    def get_loglik_comps_w(self, slot):
        o = np.zeros(3)
        ok = lib().ref_build(self.h, slot, _pd(o))
        return bool(ok), float(o[0]), float(o[1])

    def get_loglik_w(self, slot=0):
        o = np.zeros(2)
        lib().ref_loglik_w(self.h, slot, _pd(o))
        return float(o[0]), float(o[1])

    def deal_with_w(self, z=None):
        zz = None if z is None else _f(z)
        if not lib().ref_gibbs(self.h, _pd(zz)):
            raise RuntimeError("Error at gibbs_sample_w")

    def accept_make_change(self):
        lib().ref_swap(self.h)

    def predict(self, theta_changed=True):
        lib().ref_predict(self.h, int(bool(theta_changed)))

    def gibbs_sample_beta(self, zb=None):
        z = None if zb is None else _cm(zb)
        lib().ref_sample_beta(self.h, _pd(z))

    def gibbs_sample_tausq(self, fixed=None):
        f = None if fixed is None else _f(fixed)
        lib().ref_sample_tausq(self.h, _pd(f))

    @property
    def w(self):
        o = np.zeros(self.n_all)
        lib().ref_get_w(self.h, _pd(o))
        return o

    @w.setter
    def w(self, v):
        a = _f(v)
        lib().ref_set_w(self.h, _pd(a))

    def set_tausq_inv(self, t):
        a = _f(t)
        lib().ref_set_tausq_inv(self.h, _pd(a))

    def params(self):
        B, t, xb = np.zeros(self.p * self.q), np.zeros(self.q), np.zeros(self.n_all)
        lib().ref_get_params(self.h, _pd(B), _pd(t), _pd(xb))
        return {"Bcoeff": B.reshape(self.q, self.p).T.copy(), "tausq_inv": t, "XB": xb}

    def get(self, name, u=0, slot=0, c=0):
        n = lib().ref_get(self.h, name.encode(), slot, int(u), int(c), None, 0)
        if n < 0:
            raise KeyError(name)
        o = np.zeros(max(n, 1))
        lib().ref_get(self.h, name.encode(), slot, int(u), int(c), _pd(o), o.size)
        return o[:n]

    def geti(self, name, u=0, slot=0, c=0):
        return self.get(name, u, slot, c).astype(np.int64)


def kthresholds(x, k):
    x = _f(x)
    res = np.zeros(max(k - 1, 0))
    lib().ref_kthresholds(_pd(x), x.size, k, _pd(res))
    return res


def cross_covariance_ag10(coords1, mv1, coords2, mv2, ai1, ai2, phi_i, thetamv, Dmat):
    c1, c2 = np.asarray(coords1, dtype=np.float64), np.asarray(coords2, dtype=np.float64)
    n1, n2 = c1.shape[0], c2.shape[0]
    Dm = np.asarray(Dmat, dtype=np.float64)
    q = Dm.shape[1]
    a, b, m1, m2 = _cm(c1), _cm(c2), _i(mv1), _i(mv2)
    A1, A2, PH, TM, DD = _f(ai1), _f(ai2), _f(phi_i), _f(np.atleast_1d(thetamv)), _cm(Dm)
    out = np.zeros(n1 * n2)
    lib().ref_cross_covariance_ag10(_pd(a), _pi(m1), n1, _pd(b), _pi(m2), n2, _pd(A1), _pd(A2), _pd(PH), _pd(TM), TM.size, _pd(DD), q, _pd(out))
    return out.reshape(n2, n1).T.copy()


def number_revalue(original_mat, from_val, to_val):
    om = np.asarray(original_mat, dtype=np.int64)
    nr, nc = om.shape
    flat, fv, tv = _i(om.T.reshape(-1)), _i(from_val), _i(to_val)
    out = np.zeros(nr * nc, dtype=np.int64)
    lib().ref_number_revalue(_pi(flat), nr, nc, _pi(fv), _pi(tv), fv.size, _pi(out))
    return out.reshape(nc, nr).T.copy()
