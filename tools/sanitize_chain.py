"""Small pass over every kernel and launch mode of the library for compute-sanitizer (no oracle, product calls only):
   compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_chain.py [quick]
Covers the host-driven operations (both slots, eager and deferred backward half, H kept or lean, full and limited trees,
prediction) and the device-resident chain (CUDA graph, second-stream overlap of the early BUILD levels, saves on the copy
stream).  Env ST_EARLY_LEVELS / ST_GRAPH select the launch structure like in the library."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spamtree_b200 as sb  # noqa: E402
from spamtree_b200 import synth  # noqa: E402


def model(q, n, limited, **kw):
    d = synth.make_data(q, n, missing=0.1, seed=5)
    tree = sb.make_tree(d["coords"], d["y"], d["mv_id"])
    csr = (tree["indexing_ptr"], tree["indexing_idx"], tree["parents_ptr"], tree["parents_idx"], tree["children_ptr"], tree["children_idx"])
    if limited:
        csr = csr[:2] + sb.limited_edges_csr(tree, d["y"])
    th = synth.theta_for(q)
    gm = sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], tree["res_is_ref"], None, None, limited, tree["block_names"],
                       tree["block_groups"], None, np.zeros(3), th, 0.1, csr=csr, **kw)
    return gm, th


def host_driven(q, n, limited, keep_H):
    gm, th = model(q, n, limited, keep_H=keep_H)
    rng = np.random.default_rng(0)
    gm.w = rng.standard_normal(n) * .3
    print("host-driven", q, n, limited, keep_H, gm.get_loglik_comps_w(0), flush=True)
    gm.deal_with_w(rng.standard_normal(n))
    gm.get_loglik_w(0)
    gm.theta_update(1, th * 1.01)
    gm.get_loglik_comps_w(1)
    gm.accept_make_change()          # deferred half of the childless level (lean handles) + Gram refresh at the next sweep
    gm.deal_with_w(None, seed=3)     # device normals
    gm.predict(True)
    gm.gibbs_sample_tausq()
    gm.gibbs_sample_beta(None, not limited)
    gm.close()


def device_chain(q, n, limited, iters):
    gm, th = model(q, n, limited, keep_H=False)
    npar = th.size
    r = gm.mcmc(synth.default_bounds(q), np.eye(npar) * 1e-4, keep=iters, burn=iters, thin=1, rng_mode=1, seed=11,
                faithful_beta_index=not limited)
    assert np.all(np.isfinite(r["w_mcmc"])) and np.all(np.isfinite(r["yhat_mcmc"])) and np.all(np.isfinite(r["theta_mcmc"]))
    print("device chain", q, n, limited, "accepted", r["n_accepted"], "of", 2 * iters, flush=True)
    gm.close()


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    host_driven(3, 1500, False, False)
    if not quick:
        host_driven(3, 1500, False, True)
        host_driven(1, 700, True, False)
    device_chain(2, 1500, False, 3 if quick else 6)
    if not quick:
        device_chain(1, 700, True, 4)
    print("done")
