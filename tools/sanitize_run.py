"""Small end-to-end pass over every kernel mode, for compute-sanitizer:
   compute-sanitizer --tool memcheck|racecheck|initcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402


def run(q, n, limited, keep_H):
    pb = common.make_problem(q, n, limited=limited)
    gm = common.product_model(pb, keep_H=keep_H)
    rng = np.random.default_rng(0)
    gm.w = rng.standard_normal(n) * .3
    print(q, n, limited, keep_H, gm.get_loglik_comps_w(0), flush=True)
    gm.deal_with_w(rng.standard_normal(n))
    gm.get_loglik_w(0)
    gm.theta_update(1, pb["theta"] * 1.01)
    gm.get_loglik_comps_w(1)
    gm.accept_make_change()          # deferred half of the childless level (keep_H = 0) + Gram refresh at the next sweep
    gm.deal_with_w(rng.standard_normal(n))
    gm.predict(True)
    gm.gibbs_sample_tausq()
    gm.gibbs_sample_beta(None, not limited)
    gm.close()


if __name__ == "__main__":
    run(3, 1500, False, False)
    run(3, 1500, False, True)
    run(1, 700, True, False)
    print("done")
