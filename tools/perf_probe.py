"""Per-phase CUDA-event timing of the hot-path iteration at a BASELINE config. Usage: perf_probe.py C4 [iters] [n]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spamtree_b200 as sb  # noqa: E402
from spamtree_b200 import synth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    n = int(sys.argv[3]) if len(sys.argv) > 3 else None
    t0 = time.time()
    d = synth.make_config(name, n)
    q = d["q"]
    t1 = time.time()
    tree = sb.make_tree(d["coords"], d["y"], d["mv_id"])
    t2 = time.time()
    csr = (tree["indexing_ptr"], tree["indexing_idx"], tree["parents_ptr"], tree["parents_idx"], tree["children_ptr"], tree["children_idx"])
    theta = synth.theta_for(q)
    gm = sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], tree["res_is_ref"], None, None, False, tree["block_names"],
                       tree["block_groups"], None, np.zeros(3), theta, 0.1, csr=csr, keep_H=False,
                       smem_panel_bytes=int(os.environ.get("ST_SMEM_BUDGET", "0")))
    t3 = time.time()
    print(f"{name}: n={d['y'].size} q={q} blocks={tree['n_blocks']} data {t1 - t0:.1f}s tree {t2 - t1:.1f}s create {t3 - t2:.1f}s", flush=True)
    print("initial builds", gm.get_loglik_comps_w(0), gm.get_loglik_comps_w(1), flush=True)
    rng = np.random.default_rng(3)
    tot = []
    for it in range(iters):
        th = theta * (1 + 0.002 * rng.standard_normal(theta.size))
        ts = time.time()
        o, ms = gm.bench_iteration(th, do_swap=(it % 4 == 3), seed=it)
        gm.sync()
        wall = (time.time() - ts) * 1e3
        tot.append(wall)
        print(f"  it {it}: gibbs {ms[0]:.3f} llw {ms[1]:.3f} build {ms[2]:.3f} rest {ms[3]:.3f} total {ms[4]:.3f} ms | wall {wall:.3f} ms | ok {o[2]} ll {o[0]:.6g}", flush=True)
    c = gm.counters()
    best = min(tot)
    print(f"F_alg {c['f_alg']:.3e} F_exec {c['f_exec']:.3e} n_cov {c['n_cov']:.3e}; best wall {best:.3f} ms -> {1e3 / best:.1f} it/s; "
          f"F_alg rate {c['f_alg'] / best / 1e9:.2f} TF/s, executed-estimate {c['f_exec'] / best / 1e9:.2f} TF/s")


if __name__ == "__main__":
    main()
