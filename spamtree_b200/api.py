"""Host-side mirror of the reference's public interface for the hot path.

R is not available in this environment, so the R-level entry point `spamtree()` (R/spamtree_fit.R:1-371) and the Rcpp
export `spamtree_mv_mcmc` (src/spamtree_fit.cpp:5-54) are mirrored here in Python with the same argument names,
defaults and return keys; the model layer `SpamTreeMV` (src/spamtree_model.h:22-212) is a thin wrapper over the C ABI.
All compute happens in libspamtree_b200.so (CUDA, sm_100a); nothing here falls back to the CPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib

_ERRORS = {1: "invalid argument", 2: "CUDA error", 3: "not positive definite", 4: "unsupported", 5: "NaN log-likelihood"}


class SpamTreeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{_ERRORS.get(code, code)}] {msg}")
        self.code = code


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _dp(a):
    return None if a is None else a.ctypes.data_as(_lib.c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_lib.c_int64_p)


def _colmajor(a):
    """column-major (Fortran) flattening, as arma::mat memory"""
    return _f64(np.asarray(a, dtype=np.float64).T).reshape(-1) if np.ndim(a) == 2 else _f64(a).reshape(-1)


def lists_to_csr(lists):
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        ptr[i + 1] = ptr[i] + len(l)
    idx = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) if len(lists) and ptr[-1] > 0 else np.zeros(0, np.int64)
    return ptr, _i64(idx)


def csr_to_lists(ptr, idx):
    return [np.asarray(idx[ptr[i]:ptr[i + 1]], dtype=np.int64) for i in range(len(ptr) - 1)]


# --------------------------------------------------------------------------------------------- tree_dep.cpp exports
def kthresholds(x, k):
    """tree_dep.cpp:16-27"""
    x = _f64(x)
    res = np.zeros(max(k - 1, 0))
    rc = lib.st_kthresholds(_dp(x), x.size, int(k), _dp(res))
    if rc:
        raise SpamTreeError(rc, "st_kthresholds")
    return res


def part_axis_parallel_lmt(coords, thresholds):
    """tree_dep.cpp:58-67; thresholds: list (one array per axis)"""
    coords = np.asarray(coords, dtype=np.float64)
    n, d = coords.shape
    ptr, _ = lists_to_csr(thresholds)
    thr = _f64(np.concatenate([np.asarray(t, dtype=np.float64) for t in thresholds]) if ptr[-1] else np.zeros(0))
    cm = _colmajor(coords)
    out = np.zeros(n * d)
    rc = lib.st_part_axis_parallel_lmt(_dp(cm), n, d, _dp(thr), _ip(ptr), _dp(out))
    if rc:
        raise SpamTreeError(rc, "st_part_axis_parallel_lmt")
    return out.reshape(d, n).T.copy()


def number_revalue(original_mat, from_val, to_val):
    """tree_dep.cpp:240-259"""
    om = np.asarray(original_mat, dtype=np.int64)
    nr, nc = om.shape
    flat = _i64(om.T.reshape(-1))
    fv, tv = _i64(from_val), _i64(to_val)
    out = np.zeros(nr * nc, dtype=np.int64)
    rc = lib.st_number_revalue(_ip(flat), nr, nc, _ip(fv), _ip(tv), fv.size, _ip(out))
    if rc:
        raise SpamTreeError(rc, "st_number_revalue")
    return out.reshape(nc, nr).T.copy()


def _make_edges(parchimat, non_empty_blocks, res_is_ref, limited):
    pm = np.asarray(parchimat, dtype=np.float64)
    nr, L = pm.shape
    flat = _colmajor(pm)
    ne, rr = _i64(non_empty_blocks), _i64(res_is_ref)
    counts = np.zeros(3, dtype=np.int64)
    rc = lib.st_make_edges(_dp(flat), nr, L, _ip(ne), ne.size, _ip(rr), int(limited), None, None, None, None, _ip(counts))
    if rc:
        raise SpamTreeError(rc, "st_make_edges")
    nb = int(counts[0])
    pp, pi = np.zeros(nb + 1, np.int64), np.zeros(max(int(counts[1]), 1), np.int64)
    cp, ci = np.zeros(nb + 1, np.int64), np.zeros(max(int(counts[2]), 1), np.int64)
    rc = lib.st_make_edges(_dp(flat), nr, L, _ip(ne), ne.size, _ip(rr), int(limited), _ip(pp), _ip(pi), _ip(cp), _ip(ci), _ip(counts))
    if rc:
        raise SpamTreeError(rc, "st_make_edges")
    return {"parents": csr_to_lists(pp, pi), "children": csr_to_lists(cp, ci)}


def make_edges(parchimat, non_empty_blocks, res_is_ref):
    """tree_dep.cpp:75-130 (parchimat: NaN = NA, 1-based block names; returned ids are 0-based like the reference)"""
    return _make_edges(parchimat, non_empty_blocks, res_is_ref, False)


def make_edges_limited(parchimat, non_empty_blocks, res_is_ref):
    """tree_dep.cpp:133-186"""
    return _make_edges(parchimat, non_empty_blocks, res_is_ref, True)


def make_tree(coords, y, mv_id, cell_size=25, K=(2, 2), start_level=0, tree_depth=np.inf, last_not_reference=True,
              cherrypick_same_margin=True, cherrypick_group_locations=True, seed=0):
    """Deterministic stand-in for R's make_tree() + the graph-building tail of spamtree() (R/make_tree.R:1-420,
    R/spamtree_fit.R:288-324).  `coords` must be sorted by (Var1, Var2).  Returns the C++-boundary inputs."""
    coords = np.asarray(coords, dtype=np.float64)
    n = coords.shape[0]
    cm, yy, mv = _colmajor(coords), _f64(y).reshape(-1), _i64(mv_id).reshape(-1)
    o = _lib.StTreeOpts()
    o.n_all = n
    o.coords, o.y, o.mv_id = _dp(cm), _dp(yy), _ip(mv)
    o.cell_size = int(cell_size)
    o.K[0], o.K[1] = int(K[0]), int(K[1])
    o.start_level = int(start_level)
    o.tree_depth = 0 if not np.isfinite(tree_depth) else int(tree_depth)
    o.last_not_reference = int(bool(last_not_reference))
    o.cherrypick_same_margin = int(bool(cherrypick_same_margin))
    o.cherrypick_group_locations = int(bool(cherrypick_group_locations))
    o.seed = int(seed)
    t = C.c_void_p()
    rc = lib.st_make_tree(C.byref(o), C.byref(t))
    if rc:
        raise SpamTreeError(rc, lib.st_last_error(None).decode())
    try:
        sz = np.zeros(6, dtype=np.int64)
        lib.st_tree_sizes(t, _ip(sz))
        nb, nres, pr, pc, npar, nchi = [int(v) for v in sz]
        blocking, res = np.zeros(n, np.int64), np.zeros(n, np.int64)
        res_is_ref = np.zeros(nres, np.int64)
        parchi = np.zeros(pr * pc)
        ip_, ii = np.zeros(nb + 1, np.int64), np.zeros(n, np.int64)
        pp, pi = np.zeros(nb + 1, np.int64), np.zeros(max(npar, 1), np.int64)
        cp, ci = np.zeros(nb + 1, np.int64), np.zeros(max(nchi, 1), np.int64)
        bn, bg = np.zeros(nb), np.zeros(nb)
        lib.st_tree_get(t, _ip(blocking), _ip(res), _ip(res_is_ref), _dp(parchi), _ip(ip_), _ip(ii), _ip(pp), _ip(pi),
                        _ip(cp), _ip(ci), _dp(bn), _dp(bg))
    finally:
        lib.st_tree_destroy(t)
    return {
        "blocking": blocking, "res": res, "res_is_ref": res_is_ref,
        "parchi_map": parchi.reshape(pc, pr).T.copy(),
        "indexing_ptr": ip_, "indexing_idx": ii, "parents_ptr": pp, "parents_idx": pi[:npar],
        "children_ptr": cp, "children_idx": ci[:nchi], "block_names": bn, "block_groups": bg, "n_blocks": nb,
    }


def limited_edges_csr(tree, y):
    """(parents_ptr, parents_idx, children_ptr, children_idx) of make_edges_limited (tree_dep.cpp:133-186) for a tree
    returned by make_tree; y decides which blocks are non-empty (R/spamtree_fit.R:299-303)"""
    ne = np.unique(tree["blocking"][np.isfinite(np.asarray(y, dtype=np.float64).reshape(-1))])
    e = make_edges_limited(tree["parchi_map"], ne, tree["res_is_ref"])
    out = []
    for lists in (e["parents"], e["children"]):
        ptr = np.zeros(len(lists) + 1, dtype=np.int64)
        ptr[1:] = np.cumsum([len(a) for a in lists])
        idx = np.concatenate([np.asarray(a, dtype=np.int64) for a in lists]) if ptr[-1] else np.zeros(0, dtype=np.int64)
        out += [ptr, idx]
    return tuple(out)


def CrossCovarianceAG10(coords1, mv1, coords2, mv2, ai1, ai2, phi_i, thetamv, Dmat, device=0):
    """covariance_functions.cpp:301-355 (R export), evaluated on the GPU"""
    c1, c2 = np.asarray(coords1, dtype=np.float64), np.asarray(coords2, dtype=np.float64)
    n1, n2 = c1.shape[0], c2.shape[0]
    Dm = np.asarray(Dmat, dtype=np.float64)
    q = Dm.shape[1] if Dm.ndim == 2 else 1
    a, b, m1, m2 = _colmajor(c1), _colmajor(c2), _i64(mv1), _i64(mv2)
    A1, A2, PH, TM, DD = _f64(ai1), _f64(ai2), _f64(phi_i), _f64(np.atleast_1d(thetamv)), _colmajor(Dm)
    out = np.zeros(n1 * n2)
    rc = lib.st_cross_covariance_ag10(_dp(a), _ip(m1), n1, _dp(b), _ip(m2), n2, _dp(A1), _dp(A2), _dp(PH), _dp(TM), TM.size,
                                      _dp(DD), q, device, _dp(out))
    if rc:
        raise SpamTreeError(rc, lib.st_last_error(None).decode())
    return out.reshape(n2, n1).T.copy()


# --------------------------------------------------------------------------------------------- mh_adapt.h / mh_adapt.cpp
def par_huvtransf_fwd(par, set_unif_bounds):
    """mh_adapt.cpp:3-8"""
    a, b = _f64(par), _colmajor(np.asarray(set_unif_bounds, dtype=np.float64))
    out = np.zeros(a.size)
    rc = lib.st_par_huvtransf_fwd(_dp(a), a.size, _dp(b), _dp(out))
    if rc:
        raise SpamTreeError(rc, "st_par_huvtransf_fwd")
    return out


def par_huvtransf_back(par, set_unif_bounds):
    """mh_adapt.cpp:10-15"""
    a, b = _f64(par), _colmajor(np.asarray(set_unif_bounds, dtype=np.float64))
    out = np.zeros(a.size)
    rc = lib.st_par_huvtransf_back(_dp(a), a.size, _dp(b), _dp(out))
    if rc:
        raise SpamTreeError(rc, "st_par_huvtransf_back")
    return out


def mh_propose(param, set_unif_bounds, paramsd, U):
    """spamtree_fit.cpp:211-215 + calc_jacobian (mh_adapt.h:230-239): (new_param, jacobian, out_unif_bounds)"""
    a, b, sd, u = _f64(param), _colmajor(np.asarray(set_unif_bounds, dtype=np.float64)), _colmajor(np.asarray(paramsd, dtype=np.float64)), _f64(U)
    out = np.zeros(a.size + 2)
    rc = lib.st_mh_propose(a.size, _dp(a), _dp(b), _dp(sd), _dp(u), _dp(out))
    if rc:
        raise SpamTreeError(rc, "st_mh_propose")
    return out[:a.size].copy(), float(out[a.size]), bool(out[a.size + 1])


def do_I_accept(logaccept, u):
    """mh_adapt.h:20-36 with the caller's uniform draw"""
    return bool(lib.st_do_i_accept(float(logaccept), float(u)))


def ram_adapt(metropolis_sd, U, alpha):
    """class RAMAdapt (mh_adapt.h:40-135) over a recorded sequence: U steps x npar, alpha (steps).
    Returns (paramsd after the last step, trace steps x npar x npar)."""
    U = np.asarray(U, dtype=np.float64)
    steps, npar = U.shape
    sd, uu, al = _colmajor(np.asarray(metropolis_sd, dtype=np.float64)), _f64(U.reshape(-1)), _f64(alpha)
    out, tr = np.zeros(npar * npar), np.zeros(max(steps, 1) * npar * npar)
    rc = lib.st_ram_adapt(npar, _dp(sd), steps, _dp(uu), _dp(al), _dp(out), _dp(tr))
    if rc:
        raise SpamTreeError(rc, "st_ram_adapt: metropolis_sd is not positive definite")
    return out.reshape(npar, npar).T.copy(), tr[:steps * npar * npar].reshape(steps, npar, npar).transpose(0, 2, 1).copy()


# --------------------------------------------------------------------------------------------- the model layer
class SpamTreeMV:
    """src/spamtree_model.h:22-212.  Constructor arguments follow spamtree_model.cpp:8-37 (lists are 0-based id lists)."""

    def __init__(self, y, X, coords, mv_id, res_is_ref, parents, children, limited_tree, block_names, block_groups,
                 indexing, beta, theta, tausq, device=0, keep_H=True, smem_panel_bytes=0, csr=None, partition=None,
                 q=None):
        self.y = _f64(y).reshape(-1)
        self.n_all = self.y.size
        Xa = np.asarray(X, dtype=np.float64).reshape(self.n_all, -1)
        self.p = Xa.shape[1]
        self.mv_id = _i64(mv_id).reshape(-1)
        self.q = int(np.unique(self.mv_id).size) if q is None else int(q)
        self._X, self._coords = _colmajor(Xa), _colmajor(np.asarray(coords, dtype=np.float64))
        if csr is not None:
            ip_, ii, pp, pi, cp, ci = [_i64(a) for a in csr]
        else:
            ip_, ii = lists_to_csr(indexing)
            pp, pi = lists_to_csr(parents)
            cp, ci = lists_to_csr(children)
        self._csr = (ip_, ii, pp, pi, cp, ci)
        self.n_blocks = ip_.size - 1
        self._bn, self._bg, self._rr = _f64(block_names), _f64(block_groups), _i64(res_is_ref)
        self._theta, self._beta = _f64(theta), _f64(beta)
        self.npar = self._theta.size
        pr = _lib.StProblem()
        pr.n_all, pr.p, pr.q = self.n_all, self.p, self.q
        pr.y, pr.X, pr.coords, pr.mv_id = _dp(self.y), _dp(self._X), _dp(self._coords), _ip(self.mv_id)
        pr.n_blocks = self.n_blocks
        pr.indexing_ptr, pr.indexing_idx = _ip(ip_), _ip(ii)
        pr.parents_ptr, pr.parents_idx = _ip(pp), _ip(pi)
        pr.children_ptr, pr.children_idx = _ip(cp), _ip(ci)
        pr.block_names, pr.block_groups = _dp(self._bn), _dp(self._bg)
        pr.res_is_ref, pr.n_res = _ip(self._rr), self._rr.size
        pr.limited_tree = int(bool(limited_tree))
        pr.theta, pr.n_theta = _dp(self._theta), self.npar
        pr.beta, pr.tausq = _dp(self._beta), float(tausq)
        pr.device, pr.keep_H, pr.smem_panel_bytes = int(device), int(bool(keep_H)), int(smem_panel_bytes)
        if partition is not None and partition["nranks"] > 1:
            # partition: a sub-problem dict from spamtree_b200.partition.subproblem plus "allreduce": callable(ptr, count)
            fn = partition["allreduce"]

            def _cb(ctx, ptr, count):
                try:
                    fn(ptr, count)
                    return 0
                except Exception as ex:  # never let an exception cross the C boundary
                    print(f"spamtree_b200: allreduce callback failed: {ex}", flush=True)
                    return 1

            self._cb = _lib.ALLREDUCE_FN(_cb)
            self._grows = _i64(partition["global_rows"])
            pt = _lib.StPartition()
            pt.rank, pt.nranks, pt.n_top_levels = int(partition["rank"]), int(partition["nranks"]), int(partition["n_top_levels"])
            pt.rng_row_offset, pt.n_global_rows = int(partition["rng_row_offset"]), int(partition["n_global_rows"])
            pt.global_rows, pt.allreduce, pt.ctx = _ip(self._grows), self._cb, None
            self._pt = pt
            pr.partition = C.pointer(pt)
        h = C.c_void_p()
        rc = lib.st_create(C.byref(pr), C.byref(h))
        if rc:
            raise SpamTreeError(rc, lib.st_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib.st_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _chk(self, rc):
        if rc:
            raise SpamTreeError(rc, lib.st_last_error(self._h).decode())

    # -- reference-named operations (slot 0 = param_data, 1 = alter_data)
    def theta_update(self, slot, theta):
        t = _f64(theta)
        self._chk(lib.st_theta_update(self._h, slot, _dp(t)))

    def get_loglik_comps_w(self, slot):
        """returns (acceptable, loglik_w, logdetCi)"""
        o = np.zeros(3)
        self._chk(lib.st_get_loglik_comps_w(self._h, slot, _dp(o)))
        return bool(o[2]), float(o[0]), float(o[1])

    def deal_with_w(self, z=None, seed=0):
        zz = None if z is None else _f64(z).reshape(-1)
        self._chk(lib.st_deal_with_w(self._h, _dp(zz), int(seed)))

    def get_loglik_w(self, slot=0):
        o = np.zeros(2)
        self._chk(lib.st_get_loglik_w(self._h, slot, _dp(o)))
        return float(o[0]), float(o[1])

    def accept_make_change(self):
        self._chk(lib.st_accept_make_change(self._h))

    def predict(self, theta_changed=True):
        self._chk(lib.st_predict(self._h, int(bool(theta_changed))))

    def gibbs_sample_beta(self, zb=None, faithful_index=True):
        z = None if zb is None else _colmajor(np.asarray(zb, dtype=np.float64))
        self._chk(lib.st_gibbs_sample_beta(self._h, _dp(z), int(bool(faithful_index))))

    def gibbs_sample_tausq(self, fixed=None):
        f = None if fixed is None else _f64(fixed)
        self._chk(lib.st_gibbs_sample_tausq(self._h, _dp(f)))

    def seed(self, s):
        self._chk(lib.st_seed(self._h, int(s)))

    # -- public fields
    @property
    def w(self):
        o = np.zeros(self.n_all)
        self._chk(lib.st_get_w(self._h, _dp(o)))
        return o

    @w.setter
    def w(self, v):
        a = _f64(v).reshape(-1)
        self._chk(lib.st_set_w(self._h, _dp(a)))

    def params(self):
        B, t, xb = np.zeros(self.p * self.q), np.zeros(self.q), np.zeros(self.n_all)
        self._chk(lib.st_get_params(self._h, _dp(B), _dp(t), _dp(xb)))
        return {"Bcoeff": B.reshape(self.q, self.p).T.copy(), "tausq_inv": t, "XB": xb}

    def set_tausq_inv(self, t):
        a = _f64(t)
        self._chk(lib.st_set_tausq_inv(self._h, _dp(a)))

    def node_state(self, which, u=0, slot=0):
        cnt = C.c_int64(0)
        self._chk(lib.st_get_node_state(self._h, slot, int(u), which.encode(), None, 0, C.byref(cnt)))
        o = np.zeros(max(cnt.value, 1))
        self._chk(lib.st_get_node_state(self._h, slot, int(u), which.encode(), _dp(o), o.size, C.byref(cnt)))
        return o[:cnt.value]

    def index(self, which, u=0, c=0):
        cnt = C.c_int64(0)
        self._chk(lib.st_get_index(self._h, which.encode(), int(u), int(c), None, 0, C.byref(cnt)))
        o = np.zeros(max(cnt.value, 1), dtype=np.int64)
        self._chk(lib.st_get_index(self._h, which.encode(), int(u), int(c), _ip(o), o.size, C.byref(cnt)))
        return o[:cnt.value]

    def bench_iteration(self, theta_prop, do_swap=False, seed=0):
        t, o, ms = _f64(theta_prop), np.zeros(3), np.zeros(6, dtype=np.float32)
        self._chk(lib.st_bench_iteration(self._h, _dp(t), int(bool(do_swap)), int(seed), _dp(o), ms.ctypes.data_as(_lib.c_float_p)))
        return o, ms

    def set_beta_index(self, faithful):
        """row index of the beta step in bench_iteration: the reference's (App. D #12) or the corrected one"""
        self._chk(lib.st_set_beta_index(self._h, int(bool(faithful))))

    def attach_nccl(self, unique_id):
        """collective over the ranks of the partition: the library creates its own NCCL communicator from the 128-byte id of
        st_nccl_unique_id and from then on enqueues its all-reduces on its stream (no callback, no host synchronisation)"""
        b = bytes(unique_id)
        assert len(b) == 128
        self._chk(lib.st_attach_nccl(self._h, b))

    def counters(self):
        o = np.zeros(8)
        self._chk(lib.st_get_counters(self._h, _dp(o)))
        return {"launches": o[0], "f_alg": o[1], "f_exec": o[2], "n_cov": o[3], "f_alg_build": o[4], "f_exec_build": o[5],
                "b_alg_build": o[6], "chain_state_bytes": o[7]}

    def sync(self):
        self._chk(lib.st_sync(self._h))

    def mcmc(self, set_unif_bounds, mcmcsd, keep, burn, thin, adapting=True, sample_beta=True, sample_tausq=True,
             sample_theta=True, sample_w=True, sample_predicts=True, faithful_beta_index=True, rng_mode=0, seed=1,
             save_w=True, save_yhat=True):
        """the loop of spamtree_mv_mcmc (spamtree_fit.cpp:167-391)"""
        b, sd = _colmajor(np.asarray(set_unif_bounds, dtype=np.float64)), _colmajor(np.asarray(mcmcsd, dtype=np.float64))
        o = _lib.StMcmcOpts()
        o.set_unif_bounds, o.mcmcsd = _dp(b), _dp(sd)
        o.keep, o.burn, o.thin = int(keep), int(burn), int(thin)
        o.adapting, o.sample_beta, o.sample_tausq = int(adapting), int(sample_beta), int(sample_tausq)
        o.sample_theta, o.sample_w, o.sample_predicts = int(sample_theta), int(sample_w), int(sample_predicts)
        o.faithful_beta_index, o.rng_mode, o.seed = int(faithful_beta_index), int(rng_mode), int(seed)
        beta = np.zeros(self.p * keep * self.q)
        tausq, theta = np.zeros(self.q * keep), np.zeros(self.npar * keep)
        w = np.zeros(self.n_all * keep) if save_w else None
        yh = np.zeros(self.n_all * keep) if save_yhat else None
        psd = np.zeros(self.npar * self.npar)
        out = _lib.StMcmcOut()
        out.beta_mcmc, out.tausq_mcmc, out.theta_mcmc = _dp(beta), _dp(tausq), _dp(theta)
        out.w_mcmc, out.yhat_mcmc, out.paramsd = _dp(w), _dp(yh), _dp(psd)
        self._chk(lib.st_mcmc_run(self._h, C.byref(o), C.byref(out)))
        return {
            "w_mcmc": None if w is None else w.reshape(keep, self.n_all).T.copy(),
            "yhat_mcmc": None if yh is None else yh.reshape(keep, self.n_all).T.copy(),
            "beta_mcmc": beta.reshape(self.q, keep, self.p).transpose(2, 1, 0).copy(),  # p x keep x q
            "tausq_mcmc": tausq.reshape(keep, self.q).T.copy(),
            "theta_mcmc": theta.reshape(keep, self.npar).T.copy(),
            "paramsd": psd.reshape(self.npar, self.npar).T.copy(),
            "mcmc_time": out.mcmc_time, "n_accepted": out.n_accepted, "n_chol_fail": out.n_chol_fail,
        }


def spamtree_mv_mcmc(y, X, Z, coords, mv_id, blocking, gix_block, res_is_ref, parents, children, limited_tree,
                     layer_names, layer_gibbs_group, indexing, set_unif_bounds_in, start_w, theta, beta, tausq, mcmcsd,
                     mcmc_keep=100, mcmc_burn=100, mcmc_thin=1, num_threads=1, use_alg='S', adapting=False,
                     main_verbose=True, verbose=False, debug=False, printall=False, sample_beta=True,
                     sample_tausq=True, sample_theta=True, sample_w=True, sample_predicts=True, *, device=0, seed=1,
                     rng_mode=0, csr=None, save_w=True, save_yhat=True, keep_H=False):
    """Mirror of the Rcpp export (src/spamtree_fit.cpp:5-54): same positional arguments, same returned names
    (:403-414).  Z, blocking, gix_block, start_w, num_threads, use_alg and the verbosity flags are accepted and ignored
    exactly as the reference ignores them; keyword-only arguments are additions of this build."""
    model = SpamTreeMV(y, X, coords, mv_id, res_is_ref, parents, children, limited_tree, layer_names, layer_gibbs_group,
                       indexing, beta, theta, tausq, device=device, csr=csr, keep_H=keep_H)
    try:
        res = model.mcmc(set_unif_bounds_in, mcmcsd, mcmc_keep, mcmc_burn, mcmc_thin, adapting, sample_beta,
                         sample_tausq, sample_theta, sample_w, sample_predicts, True, rng_mode, seed, save_w, save_yhat)
        # the rest of the reference's returned list (spamtree_fit.cpp:410-412)
        res["block_ct_obs"] = model.index("block_ct_obs")
        res["indexing"] = indexing if indexing is not None else csr_to_lists(model._csr[0], model._csr[1])
        res["parents_indexing"] = [model.index("parents_indexing", u) for u in range(model.n_blocks)]
    finally:
        model.close()
    return res


def spamtree(y, x, coords, mv_id=None, cell_size=25, K=None, start_level=0, tree_depth=np.inf, last_not_reference=True,
             limited_tree=False, cherrypick_same_margin=True, cherrypick_group_locations=True, mvbias=0,
             mcmc=None, num_threads=4, verbose=False, settings=None, prior=None, starting=None, debug=None, *,
             device=0, seed=1, rng_mode=1, tree_seed=0):
    """Mirror of the R entry point spamtree() (R/spamtree_fit.R:1-371): same arguments and defaults, same returned
    names (`coords`, `coordsinfo`, `mv_id` + everything spamtree_mv_mcmc returns).  The tree is built by the
    deterministic stand-in for make_tree() (mvbias other than 0 is not supported)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    x = np.asarray(x, dtype=np.float64).reshape(y.size, -1)
    coords = np.asarray(coords, dtype=np.float64)
    mv_id = np.ones(y.size, dtype=np.int64) if mv_id is None else np.asarray(mv_id, dtype=np.int64)
    K = (2,) * coords.shape[1] if K is None else K
    mcmc = {"keep": 1000, "burn": 0, "thin": 1, **(mcmc or {})}
    settings = {"adapting": True, "mcmcsd": .01, "debug": False, "printall": False, **(settings or {})}
    prior = {"set_unif_bounds": None, "btmlim": None, "toplim": None, "vlim": None, **(prior or {})}
    starting = {"beta": None, "tausq": None, "theta": None, "w": None, **(starting or {})}
    debug = {"sample_beta": True, "sample_tausq": True, "sample_theta": True, "sample_w": True, "sample_predicts": True,
             **(debug or {})}
    if mvbias != 0:
        raise SpamTreeError(4, "mvbias != 0 is not supported by the deterministic tree builder")
    dd, p = coords.shape[1], x.shape[1]
    if dd > 2:
        raise SpamTreeError(4, "Not implemented in domains of dimension d>2.")  # R/spamtree_fit.R:58-60
    q = int(np.unique(mv_id).size)
    k = q * (q - 1) // 2
    start_beta = np.zeros(p) if starting["beta"] is None else np.asarray(starting["beta"], dtype=np.float64)
    btmlim = 1e-3 if prior["btmlim"] is None else prior["btmlim"]
    toplim = 1e3 if prior["toplim"] is None else prior["toplim"]
    vlim = toplim if prior["vlim"] is None else prior["vlim"]
    n_cbase = 3 if q > 2 else 1
    npars = 3 * q + n_cbase
    bounds = np.zeros((npars, 2))  # :111-134
    bounds[:, 0], bounds[:, 1] = btmlim, toplim
    if q > 1:
        bounds[1:q, 0] = -toplim
    if n_cbase == 3:
        bounds[npars - 2, :] = (btmlim, 1 - btmlim)
    if q > 1:
        vb = np.zeros((k, 2))
        vb[:, 0], vb[:, 1] = btmlim, vlim - btmlim
        bounds = np.vstack([bounds, vb])
    start_theta = bounds.mean(axis=1)  # :138
    sd = settings["mcmcsd"]
    mcmc_mh_sd = np.eye(start_theta.size) * sd if np.ndim(sd) == 0 else np.asarray(sd, dtype=np.float64)
    start_tausq = .1 if starting["tausq"] is None else starting["tausq"]
    # rows sorted by coordinates, ties by input order (:214, :267-269)
    ix = np.arange(y.size)
    order = np.lexsort((ix, coords[:, 1], coords[:, 0]))
    cs, ys, xs, mvs = coords[order], y[order], x[order], mv_id[order]
    tree = make_tree(cs, ys, mvs, cell_size, K, start_level, tree_depth, last_not_reference, cherrypick_same_margin,
                     cherrypick_group_locations, tree_seed)
    Z = np.zeros((y.size, q))
    Z[np.arange(y.size), mvs - 1] = 1
    csr = (tree["indexing_ptr"], tree["indexing_idx"], tree["parents_ptr"], tree["parents_idx"], tree["children_ptr"],
           tree["children_idx"])
    if limited_tree:  # R/spamtree_fit.R:310-311: make_edges_limited on the same parent-child map
        csr = csr[:2] + limited_edges_csr(tree, ys)
    results = spamtree_mv_mcmc(
        ys, xs, Z, cs, mvs, tree["blocking"], np.ones(y.size), tree["res_is_ref"], None, None, limited_tree,
        tree["block_names"], tree["block_groups"], None, bounds, np.zeros((y.size, q)), start_theta, start_beta,
        start_tausq, mcmc_mh_sd, mcmc["keep"], mcmc["burn"], mcmc["thin"], num_threads, 'S', settings["adapting"],
        verbose, verbose > 1, settings["debug"], settings["printall"], debug["sample_beta"], debug["sample_tausq"],
        debug["sample_theta"], debug["sample_w"], debug["sample_predicts"], device=device, seed=seed, rng_mode=rng_mode,
        csr=csr)
    coordsinfo = {"Var1": cs[:, 0], "Var2": cs[:, 1], "ix": order + 1, "block": tree["blocking"], "res": tree["res"]}
    return {"coords": cs, "coordsinfo": coordsinfo, "mv_id": mv_id, **results}
