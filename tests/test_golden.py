"""Committed known-answer vectors (tests/golden/, produced by tests/golden/make_golden.py from dense numpy math):
checked against the CPU oracle on the CPU and against the CUDA path on the GPU."""
import os

import numpy as np
import pytest

import common
from common import orc, relerr

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AG10_ARGS = ([1, 1.5], [.1, .51], [1, 2], [5.0], np.array([[0, 1.0], [1.0, 0]]))


def _problem(g):
    pb = common.make_problem(int(g["q"]), int(g["n"]))
    assert np.array_equal(pb["tree"]["blocking"], g["blocking"]), "the deterministic tree builder changed: regenerate the golden files"
    return pb


def _check_blocks(g, pb, getH, getRi, tol):
    obs_rows = np.isfinite(pb["d"]["y"])
    for key in g.files:
        if key.startswith("H_"):
            u = int(key[2:])
            H = g[key]
            if H.shape[1]:
                assert relerr(getH(u, H.shape[0]), H) <= tol, key
            Ri = g[f"Ri_{u}"]
            assert relerr(getRi(u, Ri), Ri) <= tol, f"Ri_{u}"
    return obs_rows


def test_oracle_against_golden():
    g = np.load(os.path.join(G, "ag10_manpage.npz"))
    assert relerr(orc.cross_covariance_ag10(g["coords"], g["mv"], g["coords"], g["mv"], *AG10_ARGS), g["CC"]) <= 1e-14
    g = np.load(os.path.join(G, "q3_n700.npz"))
    pb = _problem(g)
    om = common.oracle_model(pb)
    om.w = g["w"]
    ok, ll, _ = om.get_loglik_comps_w(0)
    assert ok and abs(ll - float(g["loglik"])) <= 1e-11 * abs(ll)
    isref = om.geti("block_is_reference")
    obs = _check_blocks(g, pb, lambda u, m: om.mat("H", u, m),
                        lambda u, Ri: om.mat("Ri", u, Ri.shape[0]) if isref[u] else om.get("ccholprecdiag", u), 1e-9)
    om.set_tausq_inv(g["tau"])
    om.deal_with_w(g["z"])
    assert relerr(om.w[obs], g["w_after_sweep"][obs]) <= 1e-9
    om.close()


@pytest.mark.gpu
def test_cuda_against_golden():
    import spamtree_b200 as sb
    g = np.load(os.path.join(G, "ag10_manpage.npz"))
    assert relerr(sb.CrossCovarianceAG10(g["coords"], g["mv"], g["coords"], g["mv"], *AG10_ARGS), g["CC"]) <= 1e-13
    g = np.load(os.path.join(G, "q3_n700.npz"))
    pb = _problem(g)
    gm = common.product_model(pb)
    gm.w = g["w"]
    ok, ll, _ = gm.get_loglik_comps_w(0)
    assert ok and abs(ll - float(g["loglik"])) <= 1e-10 * abs(ll)
    obs = _check_blocks(g, pb, lambda u, m: gm.node_state("H", u).reshape(-1, m).T,
                        lambda u, Ri: gm.node_state("Ri", u).reshape(Ri.shape[::-1]).T if Ri.ndim == 2 else gm.node_state("Ri", u), 1e-9)
    gm.set_tausq_inv(g["tau"])
    gm.deal_with_w(g["z"])
    assert relerr(gm.w[obs], g["w_after_sweep"][obs]) <= 1e-9
    gm.close()
