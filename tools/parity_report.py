"""Verbose parity report (product on the GPU vs the CPU oracle) — prints per-stage, per-level max relative errors.
Usage: python tools/parity_report.py [q] [n]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402


def main():
    q = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 625
    pb = common.make_problem(q, n)
    t = pb["tree"]
    nb = t["n_blocks"]
    lev = t["block_groups"].astype(int)
    m = np.diff(t["indexing_ptr"])
    t0 = time.time()
    gm = common.product_model(pb)
    print(f"q={q} n={n} blocks={nb} create {time.time() - t0:.2f}s", flush=True)
    om = common.oracle_model(pb)
    rng = np.random.default_rng(7)
    w0 = rng.standard_normal(n) * 0.5
    gm.w = w0
    om.w = w0
    for slot in (0, 1):
        rg, ro = gm.get_loglik_comps_w(slot), om.get_loglik_comps_w(slot)
        print(f"BUILD slot {slot}: gpu {rg} oracle {ro} rel {abs(rg[1] - ro[1]) / abs(ro[1]):.2e}", flush=True)
    obs = om.geti("block_ct_obs")
    isref = om.geti("block_is_reference")
    eH, eR = {}, {}
    for u in range(nb):
        if obs[u] == 0:
            continue
        P = np.diff(t["parents_ptr"])[u]
        if P > 0:
            e = common.relerr(gm.node_state("H", u), om.get("H", u))
            eH[lev[u]] = max(eH.get(lev[u], 0), e)
        a = gm.node_state("Ri", u)
        b = om.get("Ri", u) if isref[u] else om.get("ccholprecdiag", u)
        eR[lev[u]] = max(eR.get(lev[u], 0), common.relerr(a, b))
    print("H  relerr by level:", {k: f"{v:.1e}" for k, v in sorted(eH.items())})
    print("Ri relerr by level:", {k: f"{v:.1e}" for k, v in sorted(eR.items())})
    print("logdet comps relerr", common.relerr(gm.node_state("logdetCi_comps"), om.get("logdetCi_comps")),
          "loglik comps relerr", common.relerr(gm.node_state("loglik_w_comps"), om.get("loglik_w_comps")), flush=True)
    for sweep in range(2):
        z = rng.standard_normal(n)
        gm.deal_with_w(z)
        om.deal_with_w(z)
        eS, eM = {}, {}
        for u in range(nb):
            if obs[u] == 0:
                continue
            eS[lev[u]] = max(eS.get(lev[u], 0), common.relerr(gm.node_state("Sigi_tot", u), om.get("Sigi_tot", u)))
            eM[lev[u]] = max(eM.get(lev[u], 0), common.relerr(gm.node_state("Smu_tot", u), om.get("Smu_tot", u)))
        print(f"GIBBS sweep {sweep}: w relerr {common.relerr(gm.w, om.w):.2e}")
        print("   Sigi_tot by level:", {k: f"{v:.1e}" for k, v in sorted(eS.items())})
        print("   Smu_tot  by level:", {k: f"{v:.1e}" for k, v in sorted(eM.items())})
        lg, lo = gm.get_loglik_w(0), om.get_loglik_w(0)
        print(f"   LLW gpu {lg[0]:.12g} oracle {lo[0]:.12g} rel {abs(lg[0] - lo[0]) / abs(lo[0]):.2e}", flush=True)
    gm.predict(True)
    om.predict(True)
    print(f"PREDICT: w relerr {common.relerr(gm.w, om.w):.2e}")
    gm.gibbs_sample_tausq(np.full(q, 7.5))
    om.gibbs_sample_tausq(np.full(q, 7.5))
    zb = rng.standard_normal((3, q))
    gm.gibbs_sample_beta(zb, True)
    om.gibbs_sample_beta(zb)
    pg, po = gm.params(), om.params()
    print(f"BETA relerr {common.relerr(pg['Bcoeff'], po['Bcoeff']):.2e} XB relerr {common.relerr(pg['XB'], po['XB']):.2e}")
    gm.seed(5)
    om.seed(5)
    gm.gibbs_sample_tausq()
    om.gibbs_sample_tausq()
    print(f"TAUSQ relerr {common.relerr(gm.params()['tausq_inv'], om.params()['tausq_inv']):.2e}", flush=True)
    # a short lock-step chain
    from spamtree_b200 import synth
    bounds = synth.default_bounds(q)
    npar = pb["theta"].size
    gm2, om2 = common.product_model(pb), common.oracle_model(pb)
    kw = dict(keep=10, burn=20, thin=1, adapting=True, seed=11)
    rg = gm2.mcmc(bounds, np.eye(npar) * .01, rng_mode=0, **kw)
    ro = om2.mcmc(bounds, np.eye(npar) * .01, **kw)
    print(f"CHAIN(30 it): theta relerr {common.relerr(rg['theta_mcmc'], ro['theta_mcmc']):.2e} beta {common.relerr(rg['beta_mcmc'], ro['beta_mcmc']):.2e} "
          f"tausq {common.relerr(rg['tausq_mcmc'], ro['tausq_mcmc']):.2e} w {common.relerr(rg['w_mcmc'], ro['w_mcmc']):.2e} "
          f"yhat {common.relerr(rg['yhat_mcmc'], ro['yhat_mcmc']):.2e} accepted {rg['n_accepted']}/{ro['n_accepted']} gpu time {rg['mcmc_time']:.3f}s oracle {ro['mcmc_time']:.3f}s")
    print("counters", gm.counters())


if __name__ == "__main__":
    main()
