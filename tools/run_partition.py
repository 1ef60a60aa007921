"""Multi-GPU partition check, launched with torchrun (one rank per GPU): the partitioned model must reproduce the
single-GPU model — log-density, Gibbs draw given the same z, and a lock-step chain.  Usage:
   torchrun --nproc-per-node 2 tools/run_partition.py [q] [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import spamtree_b200 as sb  # noqa: E402
from spamtree_b200 import synth  # noqa: E402
from spamtree_b200 import dist as sdist  # noqa: E402
from spamtree_b200 import partition as part  # noqa: E402


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def main():
    q = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    limited = len(sys.argv) > 3 and sys.argv[3] == "limited"  # limited_tree = TRUE (make_edges_limited, tree_dep.cpp:133-186)
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    d = synth.make_data(q, n)
    tree = sb.make_tree(d["coords"], d["y"], d["mv_id"])
    theta, beta, tausq = synth.theta_for(q), np.zeros(3), 0.1
    ar = sdist.make_allreduce(torch.device("cuda", lrank))
    gm, sp, pl = sdist.partitioned_model(d, tree, theta, beta, tausq, rank, world, lrank, ar, keep_H=True, limited_tree=limited)
    csr = (tree["indexing_ptr"], tree["indexing_idx"], tree["parents_ptr"], tree["parents_idx"], tree["children_ptr"], tree["children_idx"])
    if limited:
        csr = csr[:2] + sb.limited_edges_csr(tree, d["y"])
    full = sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], tree["res_is_ref"], None, None, limited, tree["block_names"],
                         tree["block_groups"], None, beta, theta, tausq, csr=csr, device=lrank)
    rng = np.random.default_rng(3)
    w0 = rng.standard_normal(n) * .4
    gm.w = w0[sp["global_rows"]]
    full.w = w0
    ok = True
    for slot in (0, 1):
        a, b = gm.get_loglik_comps_w(slot), full.get_loglik_comps_w(slot)
        e = abs(a[1] - b[1]) / abs(b[1])
        ok &= e < 1e-12 and a[0] == b[0]
        if rank == 0:
            print(f"BUILD slot {slot}: partitioned {a[1]:.12f} single {b[1]:.12f} rel {e:.2e}", flush=True)
    for sweep in range(3):
        z = rng.standard_normal(n)
        gm.deal_with_w(z[sp["global_rows"]])
        full.deal_with_w(z)
        e = relerr(gm.w, full.w[sp["global_rows"]])
        la, lb = gm.get_loglik_w(0)[0], full.get_loglik_w(0)[0]
        ok &= e < 1e-10 and abs(la - lb) <= 1e-11 * abs(lb)
        print(f"[rank {rank}] GIBBS sweep {sweep}: w relerr {e:.2e}; LLW rel {abs(la - lb) / abs(lb):.2e}", flush=True)
    gm.predict(True)
    full.predict(True)
    e = relerr(gm.w, full.w[sp["global_rows"]])
    ok &= e < 1e-10
    for m in (gm, full):
        m.gibbs_sample_tausq(np.linspace(4, 8, q))
    zb = rng.standard_normal((3, q))
    gm.gibbs_sample_beta(zb, False)
    full.gibbs_sample_beta(zb, False)
    eb = relerr(gm.params()["Bcoeff"], full.params()["Bcoeff"])
    ok &= eb < 1e-10
    print(f"[rank {rank}] PREDICT w relerr {e:.2e}; BETA relerr {eb:.2e}", flush=True)
    # lock-step chains (host random stream drawn for every row of the problem on every rank)
    bounds = synth.default_bounds(q)
    npar = theta.size
    kw = dict(keep=8, burn=40, thin=1, adapting=True, seed=17, rng_mode=0, faithful_beta_index=False)
    sd = np.eye(npar) * (.01 if q == 1 else 1e-5)
    gm2, sp2, _ = sdist.partitioned_model(d, tree, theta, beta, tausq, rank, world, lrank, ar, limited_tree=limited)
    full2 = sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], tree["res_is_ref"], None, None, limited, tree["block_names"],
                          tree["block_groups"], None, beta, theta, tausq, csr=csr, device=lrank)
    ra = gm2.mcmc(bounds, sd, **kw)
    rb = full2.mcmc(bounds, sd, **kw)
    et, ebt = relerr(ra["theta_mcmc"], rb["theta_mcmc"]), relerr(ra["beta_mcmc"], rb["beta_mcmc"])
    ew = relerr(ra["w_mcmc"], rb["w_mcmc"][sp2["global_rows"]])
    ok &= ra["n_accepted"] == rb["n_accepted"] and et < 1e-8 and ebt < 1e-7 and ew < 1e-6
    print(f"[rank {rank}] CHAIN 48 it: accepted {ra['n_accepted']}/{rb['n_accepted']} theta {et:.2e} beta {ebt:.2e} w {ew:.2e} "
          f"time partitioned {ra['mcmc_time']:.3f}s single {rb['mcmc_time']:.3f}s", flush=True)
    # device-resident chains (rng_mode 1): Philox streams keyed by the row's id in the whole problem, so the partitioned run
    # draws what the single-GPU run draws; saves included (thin 2: saved and unsaved iterations)
    kd = dict(keep=10, burn=11, thin=2, adapting=True, seed=23, rng_mode=1, faithful_beta_index=False)
    gm3, sp3, _ = sdist.partitioned_model(d, tree, theta, beta, tausq, rank, world, lrank, ar, limited_tree=limited)
    full3 = sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], tree["res_is_ref"], None, None, limited, tree["block_names"],
                          tree["block_groups"], None, beta, theta, tausq, csr=csr, device=lrank)
    rc, rd = gm3.mcmc(bounds, sd, **kd), full3.mcmc(bounds, sd, **kd)
    et, ebt = relerr(rc["theta_mcmc"], rd["theta_mcmc"]), relerr(rc["beta_mcmc"], rd["beta_mcmc"])
    ew, ey = relerr(rc["w_mcmc"], rd["w_mcmc"][sp3["global_rows"]]), relerr(rc["yhat_mcmc"], rd["yhat_mcmc"][sp3["global_rows"]])
    ok &= rc["n_accepted"] == rd["n_accepted"] and et < 1e-8 and ebt < 1e-7 and ew < 1e-6 and ey < 1e-6
    print(f"[rank {rank}] DEVICE CHAIN 31 it: accepted {rc['n_accepted']}/{rd['n_accepted']} theta {et:.2e} beta {ebt:.2e} w {ew:.2e} yhat {ey:.2e} "
          f"time partitioned {rc['mcmc_time']:.3f}s single {rd['mcmc_time']:.3f}s", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{lrank}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PARTITION PARITY", "OK" if flag.item() == 1.0 else "FAILED", f"(gc={pl['gc']}, ranks={world}, limited_tree={limited})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
