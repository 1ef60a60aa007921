"""The drop-in boundary exercised the way its users reach it:
 (1) spamtree_b200/shim/spamtree_fit_b200.cpp — the Rcpp shim of INTEGRATION.md §1, the replacement body of the reference's
     spamtree_mv_mcmc (src/spamtree_fit.cpp:5-430) — compiled against the Armadillo/Rcpp stand-in (R is not installed here)
     and linked with libspamtree_b200.so: its returned list must equal what the ctypes path returns, bit for bit;
 (2) spamtree() / spamtree_mv_mcmc() of spamtree_b200.api, the mirrors of the R entry point (R/spamtree_fit.R:1-371), on the
     README example (README.md:30-69), with posterior summaries checked against the oracle's chain."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import common
from common import orc
import spamtree_b200 as sb
from spamtree_b200 import synth

pytestmark = pytest.mark.gpu
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)


def _shim_lib():
    subprocess.check_call(["make", "-C", os.path.join(common.ROOT, "oracle"), "shim"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(common.ROOT, "oracle", "_build", "libspamtree_shim.so"))
    L.shim_spamtree_mv_mcmc.restype = C.c_int
    L.shim_spamtree_mv_mcmc.argtypes = [C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp, _ip, C.c_int, _ip, _ip, _ip, _ip, _ip, _ip, _dp, _dp, _ip,
                                        C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_double, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_double, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _ip]
    return L


@pytest.mark.parametrize("q,n,keep,burn,thin", [(1, 625, 12, 30, 2), (3, 1500, 8, 10, 1)])
def test_rcpp_shim_equals_ctypes_path(q, n, keep, burn, thin):
    L = _shim_lib()
    pb = common.make_problem(q, n)
    d, t = pb["d"], pb["tree"]
    f, i = (lambda a: np.ascontiguousarray(a, dtype=np.float64)), (lambda a: np.ascontiguousarray(a, dtype=np.int64))
    cm = lambda a: f(np.asarray(a, dtype=np.float64).T).reshape(-1)
    y, Xc, cc, mv = f(d["y"]), cm(d["X"]), cm(d["coords"]), i(d["mv_id"])
    ip, ii, pp, pi, cp, ci = [i(a) for a in pb["csr"]]
    bn, bg, rr, th, be = f(t["block_names"]), f(t["block_groups"]), i(t["res_is_ref"]), f(pb["theta"]), f(pb["beta"])
    nb, npar, p = t["n_blocks"], th.size, 3
    bounds = synth.default_bounds(q)
    sd = np.eye(npar) * (1e-2 if q == 1 else 1e-7)
    B, S = cm(bounds), cm(sd)
    runif = 0.3141592653589793
    bm, tm, thm = np.zeros(p * keep * q), np.zeros(q * keep), np.zeros(npar * keep)
    w, yh, psd = np.zeros(n * keep), np.zeros(n * keep), np.zeros(npar * npar)
    bco, pil, pis = np.zeros(nb, np.int64), np.zeros(nb, np.int64), np.zeros(nb, np.int64)
    P = lambda a: a.ctypes.data_as(_dp)
    I = lambda a: a.ctypes.data_as(_ip)
    rc = L.shim_spamtree_mv_mcmc(n, p, q, P(y), P(Xc), P(cc), I(mv), nb, I(ip), I(ii), I(pp), I(pi), I(cp), I(ci), P(bn), P(bg), I(rr), rr.size, 0,
                                 P(th), npar, P(be), float(pb["tausq"]), P(B), P(S), keep, burn, thin, 1, 1, runif, P(bm), P(tm), P(thm), P(w),
                                 P(yh), P(psd), I(bco), I(pil), I(pis))
    assert rc == 0
    # the same run through the ctypes mirror of the Rcpp export (keep_H = 0 like the shim, the seed the shim derives from runif)
    seed = int(np.float64(runif) * np.float64(9007199254740992.0))
    r = sb.spamtree_mv_mcmc(d["y"], d["X"], np.zeros((n, q)), d["coords"], d["mv_id"], None, None, t["res_is_ref"], None, None, False,
                            t["block_names"], t["block_groups"], None, bounds, None, pb["theta"], pb["beta"], pb["tausq"], sd, keep, burn, thin,
                            adapting=True, seed=seed, rng_mode=1, csr=pb["csr"], keep_H=False)
    assert np.array_equal(thm.reshape(keep, npar).T, r["theta_mcmc"]) and np.array_equal(tm.reshape(keep, q).T, r["tausq_mcmc"])
    assert np.array_equal(bm.reshape(q, keep, p).transpose(2, 1, 0), r["beta_mcmc"])
    assert np.array_equal(w.reshape(keep, n).T, r["w_mcmc"]) and np.array_equal(yh.reshape(keep, n).T, r["yhat_mcmc"])
    assert np.array_equal(psd.reshape(npar, npar).T, r["paramsd"])
    assert np.array_equal(bco, r["block_ct_obs"])
    assert np.array_equal(pil, [a.size for a in r["parents_indexing"]]) and np.array_equal(pis, [int(a.sum()) for a in r["parents_indexing"]])
    assert np.any(np.diff(r["theta_mcmc"], axis=1) != 0)  # the chain moved


def _batch_means_se(x, nb=20):
    L = x.size // nb
    return float(x[:L * nb].reshape(nb, L).mean(axis=1).std(ddof=1) / np.sqrt(nb))


def test_spamtree_entry_point_on_the_readme_example():
    """README.md:30-69: n = 625 uniform locations, exponential covariance (sigmasq 2.3, phi 6), tausq .1, B = (-1, .5, 1), 10 %
    missing, spamtree(y - ybar, X, coords, mcmc = list(keep = 1000, burn = 1000, thin = 2)) with every default"""
    rng = np.random.default_rng(2021)
    n = 625
    coords = rng.random((n, 2))
    D = np.sqrt(((coords[:, None, :] - coords[None, :, :]) ** 2).sum(-1))
    w_latent = np.linalg.cholesky(2.3 * np.exp(-6 * D) + 1e-10 * np.eye(n)) @ rng.standard_normal(n)
    X = rng.standard_normal((n, 3))
    y_full = X @ np.array([-1, .5, 1]) + w_latent + np.sqrt(.1) * rng.standard_normal(n)
    miss = rng.random(n) < .1
    y = np.where(miss, np.nan, y_full)
    ybar = np.nanmean(y)
    res = sb.spamtree(y - ybar, X, coords, mcmc=dict(keep=1000, burn=1000, thin=2), num_threads=10, seed=7)
    # ---- the returned list: the names of spamtree_fit.cpp:403-414 plus coords / coordsinfo / mv_id (R/spamtree_fit.R:365-370)
    for k in ("coords", "coordsinfo", "mv_id", "w_mcmc", "yhat_mcmc", "beta_mcmc", "tausq_mcmc", "theta_mcmc", "paramsd", "block_ct_obs",
              "indexing", "parents_indexing", "mcmc_time"):
        assert k in res and res[k] is not None, k
    assert res["w_mcmc"].shape == (n, 1000) and res["beta_mcmc"].shape == (3, 1000, 1) and res["theta_mcmc"].shape == (4, 1000)
    order = res["coordsinfo"]["ix"] - 1            # rows come back sorted by coordinates (R/spamtree_fit.R:214, :365-366)
    assert np.array_equal(res["coords"], coords[order])
    assert int(res["block_ct_obs"].sum()) == int((~miss).sum())
    assert len(res["indexing"]) == len(res["parents_indexing"]) == res["block_ct_obs"].size
    # ---- predictions as the README forms them: posterior means of yhat and w
    y_out = res["yhat_mcmc"].mean(axis=1) + ybar
    w_out = res["w_mcmc"].mean(axis=1)
    ms, os_ = miss[order], ~miss[order]
    assert np.corrcoef(y_out[os_], y_full[order][os_])[0, 1] > 0.97       # observed rows: fitted
    assert np.corrcoef(y_out[ms], y_full[order][ms])[0, 1] > 0.75         # held-out rows: predicted
    assert np.corrcoef(w_out, w_latent[order])[0, 1] > 0.8
    # ---- posterior summaries against the oracle's chain on the same inputs (its own host random stream: two estimates of
    # the same posterior, compared within Monte-Carlo error)
    cs, ys, xs = coords[order], (y - ybar)[order], X[order]
    mv = np.ones(n, dtype=np.int64)
    tree = sb.make_tree(cs, ys, mv)
    assert np.array_equal(tree["blocking"], res["coordsinfo"]["block"])
    csr = (tree["indexing_ptr"], tree["indexing_idx"], tree["parents_ptr"], tree["parents_idx"], tree["children_ptr"], tree["children_idx"])
    bounds = synth.default_bounds(1)
    om = orc.OracleModel(ys, xs, cs, mv, tree["res_is_ref"], csr, False, tree["block_names"], tree["block_groups"], np.zeros(3),
                         bounds.mean(axis=1), .1, flags=orc.FLAG_CORRECT_PREDICT_CACHE)
    ro = om.mcmc(bounds, np.eye(4) * .01, 1000, 1000, 2, seed=3)
    om.close()
    pairs = [(f"beta[{a}]", res["beta_mcmc"][a, :, 0], ro["beta_mcmc"][a, :, 0]) for a in range(3)]
    pairs += [("tausq", res["tausq_mcmc"][0], ro["tausq_mcmc"][0]), ("sigmasq", res["theta_mcmc"][0], ro["theta_mcmc"][0]),
              ("phi", res["theta_mcmc"][3], ro["theta_mcmc"][3])]
    for name, a, b in pairs:
        se = np.hypot(_batch_means_se(a), _batch_means_se(b))
        assert abs(a.mean() - b.mean()) <= 4.5 * se, (name, a.mean(), b.mean(), se)
    assert np.corrcoef(w_out, ro["w_mcmc"].mean(axis=1))[0, 1] > 0.99
