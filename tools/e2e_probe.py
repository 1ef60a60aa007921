"""Acceptance rate and time per iteration of the public MCMC driver at a BASELINE config, for several proposal scales.
Usage: e2e_probe.py C4 [iters] [sd ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spamtree_b200 as sb  # noqa: E402
from spamtree_b200 import synth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    sds = [float(x) for x in sys.argv[3:]] or [1e-4, 1e-5, 1e-6]
    d = synth.make_config(name)
    tree = sb.make_tree(d["coords"], d["y"], d["mv_id"])
    csr = (tree["indexing_ptr"], tree["indexing_idx"], tree["parents_ptr"], tree["parents_idx"], tree["children_ptr"], tree["children_idx"])
    theta = synth.theta_for(d["q"])
    gm = sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], tree["res_is_ref"], None, None, False, tree["block_names"],
                       tree["block_groups"], None, np.zeros(3), theta, 0.1, csr=csr, keep_H=False)
    bounds = synth.default_bounds(d["q"])
    npar = theta.size
    for sd in sds:
        for adapting in (False, True):
            res = gm.mcmc(bounds, np.eye(npar) * sd, keep=iters, burn=0, thin=1, adapting=adapting, rng_mode=1, seed=5,
                          sample_predicts=False, save_w=True, save_yhat=False)
            print(f"sd {sd:g} adapting {adapting}: {res['n_accepted']} of {iters} accepted, {res['n_chol_fail']} chol failures, "
                  f"{1e3 * res['mcmc_time'] / iters:.3f} ms/iteration -> {iters / res['mcmc_time']:.1f} it/s; theta[0:3] {res['theta_mcmc'][:3, -1]}", flush=True)


if __name__ == "__main__":
    main()
