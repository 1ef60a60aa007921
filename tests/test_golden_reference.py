"""Known-answer vectors produced by THE REFERENCE'S OWN CODE (tests/golden/ref_*.npz, written by
tests/golden/make_golden_from_reference.py from oracle/_ref = the reference's model layer compiled unmodified against
the Armadillo stand-in).  The CPU oracle is checked against them on the CPU, the CUDA path on the GPU — the latter is the
direct statement "same inputs -> the reference's numbers" that does not go through the oracle.
Tolerances: integer structures bit-exact; floating point 1e-9 relative (north_star), 2e-8 for q = 1 where the
reference's cexpcov takes the distance from a norm expansion whose rounding noise the product does not reproduce
(DESIGN.md §2)."""
import glob
import os

import numpy as np
import pytest

import common
from common import relerr

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(G, "ref_q*_n*.npz")))   # *_limited.npz: limited_tree = TRUE


def _tol(q):
    return 2e-8 if q == 1 else 1e-9


def _problem(g):
    pb = common.make_problem(int(g["q"]), int(g["n"]), limited=bool(int(g["limited"])) if "limited" in g.files else False)
    assert np.array_equal(pb["tree"]["blocking"], g["blocking"]), "the deterministic tree builder changed: regenerate the golden files"
    return pb


def _run(g, pb, m, getH, getRi, geti):
    """the same call sequence as the generator, on model m; returns nothing, asserts"""
    q, tol = int(g["q"]), _tol(int(g["q"]))
    nb = pb["tree"]["n_blocks"]
    for name in ["blocks_not_empty", "blocks_predicting", "block_is_reference", "block_ct_obs"]:
        assert np.array_equal(geti(name, None), g["i_" + name]), name
    for name in ["parents_indexing", "children_indexing", "dim_by_parent", "this_is_jth_child"]:
        ptr, val = g["i_" + name + "_ptr"], g["i_" + name]
        for u in range(nb):
            assert np.array_equal(geti(name, u), val[ptr[u]:ptr[u + 1]]), (name, u)
    m.w = g["w0"]
    ok, ll, ld = m.get_loglik_comps_w(0)
    assert ok and abs(ll - float(g["loglik"])) <= tol * abs(ll) and abs(ld - float(g["logdet"])) <= tol * abs(ld)
    for key in g.files:
        if key.startswith("Ri_"):
            u = int(key[3:])
            assert relerr(getRi(u), g[key]) <= tol, key
            if f"H_{u}" in g.files:
                assert relerr(getH(u), g[f"H_{u}"]) <= tol, f"H_{u}"
    obs = np.isfinite(pb["d"]["y"])
    m.set_tausq_inv(g["tau"])
    for z, wk, lk in [("z1", "w_sweep1", "llw_sweep1"), ("z2", "w_sweep2", "llw_sweep2")]:
        m.deal_with_w(g[z])
        assert relerr(m.w[obs], g[wk][obs]) <= tol, wk
        l = m.get_loglik_w(0)[0]
        assert abs(l - float(g[lk])) <= tol * abs(l), lk
    m.theta_update(1, g["theta2"])
    ok2, ll2, ld2 = m.get_loglik_comps_w(1)
    assert ok2 and abs(ll2 - float(g["loglik2"])) <= tol * abs(ll2) and abs(ld2 - float(g["logdet2"])) <= tol * abs(ld2)
    m.accept_make_change()
    m.deal_with_w(g["z3"])
    assert relerr(m.w[obs], g["w_sweep3"][obs]) <= tol
    l = m.get_loglik_w(0)[0]
    assert abs(l - float(g["llw_sweep3"])) <= tol * abs(l)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_against_reference_outputs(path):
    g = np.load(path)
    pb = _problem(g)
    om = common.oracle_model(pb)
    isref = om.geti("block_is_reference")
    _run(g, pb, om, lambda u: om.get("H", u), lambda u: om.get("Ri", u) if isref[u] else om.get("ccholprecdiag", u),
         lambda name, u: om.geti(name) if u is None else om.geti(name, u))
    om.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_cuda_against_reference_outputs(path):
    g = np.load(path)
    pb = _problem(g)
    gm = common.product_model(pb)
    _run(g, pb, gm, lambda u: gm.node_state("H", u), lambda u: gm.node_state("Ri", u),
         lambda name, u: gm.index(name) if u is None else gm.index(name, u))
    gm.close()


def test_fixtures_are_present():
    assert len(FILES) >= 4 and any(f.endswith("_limited.npz") for f in FILES)
    assert len(CHAIN_FILES) >= 3


# ---- whole chains of the reference's OWN DRIVER (spamtree_mv_mcmc, spamtree_fit.cpp:5-430, compiled unmodified) on the
# host random stream every implementation shares (rng_mode 0): proposals, Jacobian, accept decisions, RAM adaptation,
# beta / tausq draws, prediction and the saved w / yhat of a whole run
CHAIN_FILES = sorted(glob.glob(os.path.join(G, "ref_chain_q*_n*.npz")))


def _chain_check(g, r, tol_theta, tol_par, tol_rows):
    keep = int(g["keep"])
    sel = [0, keep // 2, keep - 1]
    errs = {"theta_mcmc": relerr(r["theta_mcmc"], g["theta_mcmc"]), "beta_mcmc": relerr(r["beta_mcmc"], g["beta_mcmc"]),
            "tausq_mcmc": relerr(r["tausq_mcmc"], g["tausq_mcmc"]), "paramsd": relerr(r["paramsd"], g["paramsd"]),
            "w": relerr(r["w_mcmc"][:, sel], g["w_saved"]), "yhat": relerr(r["yhat_mcmc"][:, sel], g["yhat_saved"])}
    print("chain vs the reference's driver:", {k: f"{v:.2e}" for k, v in errs.items()})
    # identical accept / reject decisions: theta changes between exactly the same saved iterations
    assert np.array_equal(np.any(np.diff(r["theta_mcmc"], axis=1) != 0, axis=0), np.any(np.diff(g["theta_mcmc"], axis=1) != 0, axis=0))
    assert errs["theta_mcmc"] <= tol_theta and errs["paramsd"] <= tol_par
    assert errs["beta_mcmc"] <= tol_par and errs["tausq_mcmc"] <= tol_par
    assert errs["w"] <= tol_rows and errs["yhat"] <= tol_rows


def _chain_args(g):
    from spamtree_b200 import synth
    q = int(g["q"])
    pb = common.make_problem(q, int(g["n"]), limited=bool(int(g["limited"])))
    assert np.array_equal(pb["tree"]["blocking"], g["blocking"]), "the deterministic tree builder changed: regenerate the golden files"
    kw = dict(keep=int(g["keep"]), burn=int(g["burn"]), thin=int(g["thin"]), adapting=bool(int(g["adapting"])),
              sample_predicts=bool(int(g["predicts"])), seed=int(g["seed"]))
    return pb, synth.default_bounds(q), np.eye(pb["theta"].size) * float(g["sd"]), kw


@pytest.mark.parametrize("path", CHAIN_FILES, ids=[os.path.basename(f) for f in CHAIN_FILES])
def test_oracle_chain_against_reference_driver_outputs(path):
    g = np.load(path)
    pb, bounds, sd, kw = _chain_args(g)
    om = common.oracle_model(pb, flags=0)
    assert np.array_equal(om.geti("block_ct_obs"), g["block_ct_obs"])
    _chain_check(g, om.mcmc(bounds, sd, **kw), 5e-9, 5e-9, 5e-9)
    om.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", CHAIN_FILES, ids=[os.path.basename(f) for f in CHAIN_FILES])
def test_cuda_chain_against_reference_driver_outputs(path):
    """st_mcmc_run in lock-step mode (rng_mode 0) vs the reference driver's chain: identical accept decisions, every saved
    draw to 1e-9 (measured on the B200: 1e-14 after up to 84 iterations)"""
    g = np.load(path)
    pb, bounds, sd, kw = _chain_args(g)
    gm = common.product_model(pb)
    assert np.array_equal(gm.index("block_ct_obs"), g["block_ct_obs"])
    plen = np.array([gm.index("parents_indexing", u).size for u in range(pb["tree"]["n_blocks"])])
    assert np.array_equal(plen, g["parents_indexing_len"])
    _chain_check(g, gm.mcmc(bounds, sd, rng_mode=0, **kw), 1e-9, 1e-9, 1e-9)
    gm.close()
