"""Size-independent properties at (or near) BASELINE sizes, where the oracle would take minutes: consistency of the
two log-density paths, run-to-run determinism, sampled blocks against dense math, swap invariance."""
import numpy as np
import pytest

import common
from common import relerr
from dense_twin import cov
from spamtree_b200 import synth
import spamtree_b200 as sb

pytestmark = pytest.mark.gpu


def _model(name, n=None, keep_H=True):
    d = synth.make_config(name, n)
    t = sb.make_tree(d["coords"], d["y"], d["mv_id"])
    csr = (t["indexing_ptr"], t["indexing_idx"], t["parents_ptr"], t["parents_idx"], t["children_ptr"], t["children_idx"])
    th = synth.theta_for(d["q"])
    gm = sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], t["res_is_ref"], None, None, False, t["block_names"],
                       t["block_groups"], None, np.zeros(3), th, 0.1, csr=csr, keep_H=keep_H)
    return d, t, th, gm


@pytest.mark.parametrize("name,n", [("C3", None), ("C4", None)])
def test_fullsize_properties(name, n):
    d, t, th, gm = _model(name, n)
    N = d["y"].size
    ok, ll0, ld0 = gm.get_loglik_comps_w(0)
    assert ok and np.isfinite(ll0)
    ok1, ll1, ld1 = gm.get_loglik_comps_w(1)
    assert (ll1, ld1) == (ll0, ld0)                      # same theta in both slots: bitwise identical (deterministic)
    rng = np.random.default_rng(0)
    z = rng.standard_normal(N)
    gm.deal_with_w(z)
    w1 = gm.w
    l_llw = gm.get_loglik_w(0)[0]
    okb, l_build, _ = gm.get_loglik_comps_w(0)           # BUILD recomputes the density from scratch at the new w
    assert okb and abs(l_llw - l_build) <= 1e-10 * abs(l_build)
    # determinism: a second model fed the same z gives bit-identical w
    gm.w = np.zeros(N)
    gm.get_loglik_comps_w(0)
    gm.deal_with_w(z)
    assert np.array_equal(gm.w, w1)
    # sampled deep blocks against dense math: H = K_{u,pa} K_pa^-1, Ri = chol(K_uu - H K_pa,u)^-1
    ip, ii, pp, pi = t["indexing_ptr"], t["indexing_idx"], t["parents_ptr"], t["parents_idx"]
    lev = t["block_groups"].astype(int)
    obs_levels = sorted(set(lev[np.array([np.isfinite(d["y"][ii[ip[u]]]) for u in range(t["n_blocks"])])]))
    deep_ref, leaf = obs_levels[-2], obs_levels[-1]
    for L, isref in ((deep_ref, True), (leaf, False)):
        us = np.flatnonzero(lev == L)
        for u in rng.choice(us, size=4, replace=False):
            ru = ii[ip[u]:ip[u + 1]]
            rp = np.concatenate([ii[ip[a]:ip[a + 1]] for a in pi[pp[u]:pp[u + 1]]])
            Kpp = cov(d["coords"][rp], d["mv_id"][rp], d["coords"][rp], d["mv_id"][rp], th, d["q"])
            Kup = cov(d["coords"][ru], d["mv_id"][ru], d["coords"][rp], d["mv_id"][rp], th, d["q"])
            Kuu = cov(d["coords"][ru], d["mv_id"][ru], d["coords"][ru], d["mv_id"][ru], th, d["q"])
            H = np.linalg.solve(Kpp, Kup.T).T
            R = Kuu - H @ Kup.T
            m = ru.size
            assert relerr(gm.node_state("H", u).reshape(-1, m).T, H) <= 1e-8
            got = gm.node_state("Ri", u)
            want = np.linalg.inv(np.linalg.cholesky((R + R.T) / 2)) if isref else 1 / np.sqrt(np.diag(R))
            assert relerr(got.reshape(m, m).T if isref else got, want) <= 1e-8
    # accept/swap: the alter slot becomes the param slot without recomputation
    th2 = th * (1 + 1e-3 * rng.standard_normal(th.size))
    gm.theta_update(1, th2)
    ok2, ll2, _ = gm.get_loglik_comps_w(1)
    assert ok2
    gm.accept_make_change()
    assert abs(gm.get_loglik_w(0)[0] - ll2) <= 1e-12 * abs(ll2)
    gm.predict(True)
    assert np.all(np.isfinite(gm.w))
    gm.close()


@pytest.mark.parametrize("name", ["C3", "C4"])
def test_fullsize_deferred_half_is_consistent(name):
    """keep_H = 0 (the bench configuration): the childless level gets the forward half of BUILD only and the backward half
    when the proposal is accepted.  Size-independent checks: the log-density of the forward half equals what LLW computes
    from the completed G after the swap; a rejected proposal does not disturb the current slot; run-to-run determinism."""
    d, t, th, gm = _model(name, keep_H=False)
    N = d["y"].size
    rng = np.random.default_rng(1)
    z = rng.standard_normal(N)
    ok, ll0, _ = gm.get_loglik_comps_w(0)
    assert ok
    gm.deal_with_w(z)                                    # completes slot 0 (deferred half) before the sweep
    w1 = gm.w
    l_llw = gm.get_loglik_w(0)[0]
    okb, l_build, _ = gm.get_loglik_comps_w(0)           # forward half again, at the new w
    assert okb and abs(l_llw - l_build) <= 1e-10 * abs(l_build)
    th2 = th * (1 + 1e-3 * rng.standard_normal(th.size))
    gm.theta_update(1, th2)
    ok2, ll2, _ = gm.get_loglik_comps_w(1)               # proposal: forward half only at the leaves
    assert ok2
    # rejected: the current slot still gives the same LLW
    assert gm.get_loglik_w(0)[0] == l_llw
    gm.accept_make_change()                              # accepted: backward half runs now
    assert abs(gm.get_loglik_w(0)[0] - ll2) <= 1e-10 * abs(ll2)
    gm.deal_with_w(z)
    w2 = gm.w
    # determinism: replay from the same state
    gm.w = w1
    gm.get_loglik_comps_w(0)
    gm.deal_with_w(z)
    assert np.array_equal(gm.w, w2)
    gm.close()
