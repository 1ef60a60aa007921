"""Pins the CPU oracle.  The reference ships no tests or golden vectors, so the oracle is pinned by
(a) known answers derived from the reference source and man pages, and (b) the dense-math twin of tests/dense_twin.py."""
import numpy as np
import pytest

import common
from common import orc, relerr
from dense_twin import Twin, cov


def test_kat_ag10_man_page_example():
    """man/CrossCovarianceAG10.Rd:66-95: h = 0 gives ai1^2 + ai2^2 on same-outcome entries and
    ai1[0]*ai1[1]/(delta+1) on cross entries (covariance_functions.cpp:123-128, 250-255)"""
    xl = np.linspace(0, 1, 10)
    g = np.array([(a, b) for b in xl for a in xl])
    coords = np.vstack([g, g])
    mv = np.r_[np.ones(100, int), 2 * np.ones(100, int)]
    ai1, ai2, phi_i, thetamv = [1, 1.5], [.1, .51], [1, 2], [5.0]
    D = np.array([[0, 1.0], [1.0, 0]])
    CC = orc.cross_covariance_ag10(coords, mv, coords, mv, ai1, ai2, phi_i, thetamv, D)
    assert CC.shape == (200, 200)
    assert np.allclose(np.diag(CC)[:100], 1.01, rtol=0, atol=1e-15)
    assert np.allclose(np.diag(CC)[100:], 2.5101, rtol=0, atol=1e-15)
    assert np.allclose(CC[np.arange(100), 100 + np.arange(100)], 0.75, rtol=0, atol=1e-15)
    assert np.allclose(CC, CC.T, rtol=0, atol=1e-15)
    # an off-diagonal entry by hand: h = 1/9 between grid neighbours, same outcome 1, cross with delta = 1
    h = 1 / 9
    assert CC[0, 1] == pytest.approx(1.0 * np.exp(-5 * h) + .01 * np.exp(-1 * h), rel=1e-14)
    assert CC[0, 101] == pytest.approx(1.5 * np.exp(-5 * h / np.sqrt(2)) / 2, rel=1e-14)
    # and against the dense twin's independent statement of the formula
    th = np.r_[ai1, ai2, phi_i, thetamv, 1.0]
    assert relerr(CC, cov(coords, mv, coords, mv, th, 2)) < 1e-14
    assert np.linalg.eigvalsh(CC).min() > 0


def test_kat_q3_psi_and_vec_to_symmat():
    """q > 2: psi = (a v + 1)^beta (covariance_functions.h:44-48); delta fill order (2,1),(3,1),(3,2) (:77-92)"""
    c = np.array([[0.1, 0.2], [0.4, 0.6], [0.9, 0.3]])
    mv = np.array([1, 2, 3])
    ai1, ai2, phi = np.array([1, 1.5, .8]), np.array([.1, .51, .3]), np.array([1, 2, 3.0])
    a, beta, cc = 2.0, .5, 5.0
    D = np.zeros((3, 3))
    D[1, 0] = D[0, 1] = 1.0
    D[2, 0] = D[0, 2] = 2.0
    D[2, 1] = D[1, 2] = 1.5
    K = orc.cross_covariance_ag10(c, mv, c, mv, ai1, ai2, phi, [a, beta, cc], D)
    for i in range(3):
        for j in range(3):
            h = np.linalg.norm(c[i] - c[j])
            if i == j:
                want = ai1[i] ** 2 + ai2[i] ** 2
            else:
                psi = (a * D[i, j] + 1) ** beta
                want = ai1[i] * ai1[j] * np.exp(-cc * h / np.sqrt(psi)) / psi
            assert K[i, j] == pytest.approx(want, rel=1e-14)


def test_kat_kthresholds():
    assert list(orc.kthresholds(np.arange(1, 101.0), 4)) == [26, 51, 76]  # tree_dep.cpp:16-27
    x = np.random.default_rng(0).random(1000)
    t = orc.kthresholds(x, 10)
    xs = np.sort(x)
    assert all(t[i - 1] == xs[i * 1000 // 10] for i in range(1, 10))


@pytest.mark.parametrize("q,n", [(1, 500), (2, 700), (3, 900)])
def test_oracle_build_matches_dense_math(q, n):
    pb = common.make_problem(q, n)
    om, tw = common.oracle_model(pb), Twin(pb)
    rng = np.random.default_rng(5)
    w = rng.standard_normal(n)
    om.w = w
    ok, ll, ld = om.get_loglik_comps_w(0)
    assert ok
    nb = pb["tree"]["n_blocks"]
    worst = 0.0
    for u in range(nb):
        if tw.obs[u] == 0:
            continue
        H, Ri = tw.block(u)
        m = tw.rows[u].size
        if H.shape[1]:
            worst = max(worst, relerr(om.mat("H", u, m), H))
        if tw.isref[u]:
            worst = max(worst, relerr(om.mat("Ri", u, m), Ri))
            # invariant (i): Kxx_inv(u) built by the block recursion == inverse of the dense chain covariance
            if om.get("Kxx_inv", u).size:
                rows = np.r_[tw.prow(u), tw.rows[u]].astype(int)
                Kd = tw.K(rows, rows)
                worst = max(worst, relerr(om.mat("Kxx_inv", u, rows.size), np.linalg.inv(Kd)))
        else:
            worst = max(worst, relerr(om.get("ccholprecdiag", u), Ri))
    assert worst < 5e-9, worst
    # invariant (ii): loglik_w == MVN log-density with precision Q
    want = tw.loglik(w)
    assert abs(ll - want) <= 1e-10 * abs(want)
    assert abs(om.get_loglik_w(0)[0] - want) <= 1e-10 * abs(want)
    om.close()


@pytest.mark.parametrize("q,n", [(1, 500), (3, 900)])
def test_oracle_gibbs_matches_dense_math_with_reference_message_timing(q, n):
    """invariant (iii) + SURVEY App. D #11: conditionals assembled from messages formed at sampling time"""
    pb = common.make_problem(q, n)
    om, tw = common.oracle_model(pb), Twin(pb)
    rng = np.random.default_rng(6)
    w = rng.standard_normal(n) * .3
    om.w = w
    assert om.get_loglik_comps_w(0)[0]
    tau = np.array([3.0, 5.0, 8.0])[:q]
    om.set_tausq_inv(tau)
    p = om.params()
    y0 = np.where(np.isfinite(pb["d"]["y"]), pb["d"]["y"], 0.0)
    resid = y0 - p["XB"]
    for _ in range(2):
        z = rng.standard_normal(n)
        om.deal_with_w(z)
        w, probes = tw.gibbs_sweep(w, z, tau[pb["d"]["mv_id"] - 1], resid)
        obs_rows = np.concatenate([tw.rows[u] for u in range(tw.nb) if tw.obs[u] > 0])
        assert relerr(om.w[obs_rows], w[obs_rows]) < 1e-9
        for u, (Sig, Smu) in probes.items():
            m = tw.rows[u].size
            got = om.mat("Sigi_tot", u, m) if tw.isref[u] else om.get("Sigi_tot", u)
            assert relerr(got, Sig) < 1e-9
            assert relerr(om.get("Smu_tot", u), Smu) < 1e-8
        w = om.w.copy()  # keep the two in lock-step (prediction rows untouched by the sweep)
    om.close()


def test_oracle_chol_failure_is_a_rejection_not_an_error():
    pb = common.make_problem(3, 900)
    om = common.oracle_model(pb)
    assert om.get_loglik_comps_w(0)[0]
    bad = pb["theta"].copy()
    bad[12:] = [1e-3, 1e-3, 999.0]  # cross-distances that make the cross-covariance indefinite
    bad[0:3] = [30, -30, 30]
    om.theta_update(1, bad)
    ok, ll, _ = om.get_loglik_comps_w(1)
    if ok:
        pytest.skip("theta did not break positive-definiteness on this tree")
    om.theta_update(1, pb["theta"])
    assert om.get_loglik_comps_w(1)[0]  # the slot recovers on the next proposal
    om.close()
