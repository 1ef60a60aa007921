"""Generates tests/golden/*.npz from the dense numpy statement of the math (tests/dense_twin.py) — independent of both
the CPU oracle and the CUDA path.  The reference itself cannot run here (no R / Armadillo), so these are the committed
known-answer vectors.  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import common  # noqa: E402
from dense_twin import Twin, cov  # noqa: E402


def main():
    # 1. man/CrossCovarianceAG10.Rd:66-95
    xl = np.linspace(0, 1, 10)
    g = np.array([(a, b) for b in xl for a in xl])
    coords = np.vstack([g, g])
    mv = np.r_[np.ones(100, int), 2 * np.ones(100, int)]
    th = np.r_[[1, 1.5], [.1, .51], [1, 2], [5.0], [1.0]]
    np.savez_compressed(os.path.join(HERE, "ag10_manpage.npz"), coords=coords, mv=mv, CC=cov(coords, mv, coords, mv, th, 2))
    # 2. a small q = 3 tree: H, Ri per block, log-density and one Gibbs sweep at fixed inputs
    pb = common.make_problem(3, 700)
    tw = Twin(pb)
    rng = np.random.default_rng(42)
    w = rng.standard_normal(700) * .5
    z = rng.standard_normal(700)
    out = {"q": 3, "n": 700, "w": w, "z": z, "theta": pb["theta"], "loglik": tw.loglik(w), "blocking": pb["tree"]["blocking"]}
    for u in range(tw.nb):
        if tw.obs[u] == 0:
            continue
        H, Ri = tw.block(u)
        out[f"H_{u}"], out[f"Ri_{u}"] = H, Ri
    tau = np.array([3.0, 5.0, 8.0])
    y0 = np.where(np.isfinite(pb["d"]["y"]), pb["d"]["y"], 0.0)
    w1, _ = tw.gibbs_sweep(w, z, tau[pb["d"]["mv_id"] - 1], y0)   # XB = 0 at beta = 0
    out["tau"], out["w_after_sweep"] = tau, w1
    np.savez_compressed(os.path.join(HERE, "q3_n700.npz"), **out)
    print("written")


if __name__ == "__main__":
    main()
