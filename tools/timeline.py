"""Kernel timeline of the device-resident chain (diagnosis only — numbers taken under a profiler are never bench values).
torch.profiler (CUPTI activity records) sees every kernel of the process, including the ones this library launches through
the CUDA runtime and its CUDA graphs.  Usage: timeline.py C1 [iterations] -> one iteration's kernels with start offset,
duration, stream and the gap to the previous kernel's end on that stream; trace in gpurun_out/timeline_<cfg>.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import spamtree_b200 as sb  # noqa: E402
from spamtree_b200 import synth  # noqa: E402


def report(path, name, head=""):
    """one accepted and one rejected iteration of a saved trace (also offline: timeline.py --json trace.json)"""
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    print(f"{name}: {head}; {len(ev)} device activities")
    # iteration boundaries: the childless level opens every sweep (the normals are drawn underneath the previous iteration's tail)
    starts = [i for i, e in enumerate(ev) if "gibbs_level_kernel<0>" in e["name"] or "gibbs_level_kernel<(int)0>" in e["name"]]
    if len(starts) < 4:
        print("no iteration markers found")
        return
    # one accepted and one rejected iteration (the Gram refresh runs only after an accepted proposal)
    def gram_us(i):
        return sum(e["dur"] for e in ev[starts[i]:starts[i + 1]] if "gram_level" in e["name"])
    mid = range(2, len(starts) - 1)
    acc = [i for i in mid if gram_us(i) > 15.0]
    rej = [i for i in mid if gram_us(i) <= 15.0]
    print("iteration lengths (us):", " ".join(f"{ev[starts[i + 1]]['ts'] - ev[starts[i]]['ts']:.0f}{'a' if i in acc else 'r'}" for i in mid))
    for which in acc[:1] + rej[:1]:
        i0, i1 = starts[which], starts[which + 1]
        t0 = ev[i0]["ts"]
        print(f"--- iteration {which} ({'accepted' if which in acc else 'rejected'}): {ev[i1]['ts'] - t0:.1f} us")
        last_end = {}
        for e in ev[i0:i1]:
            st = e.get("args", {}).get("stream", -1)
            gap = e["ts"] - last_end.get(st, e["ts"])
            last_end[st] = e["ts"] + e["dur"]
            nm = e["name"].replace("st::", "").replace("void ", "").split("(")[0][:34]
            grid = e.get("args", {}).get("grid", "")
            print(f"  +{e['ts'] - t0:7.1f}  dur {e['dur']:6.1f}  end {e['ts'] + e['dur'] - t0:7.1f}  gap {gap:6.1f}  s{st:<3} {nm:34s} {grid}")


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--json":
        report(sys.argv[2], os.path.basename(sys.argv[2]))
        return
    name = sys.argv[1] if len(sys.argv) > 1 else "C1"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    sd = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-8
    torch.cuda.init()
    d = synth.make_config(name)
    q = d["q"]
    tree = sb.make_tree(d["coords"], d["y"], d["mv_id"])
    csr = (tree["indexing_ptr"], tree["indexing_idx"], tree["parents_ptr"], tree["parents_idx"], tree["children_ptr"], tree["children_idx"])
    theta = synth.theta_for(q)
    gm = sb.SpamTreeMV(d["y"], d["X"], d["coords"], d["mv_id"], tree["res_is_ref"], None, None, False, tree["block_names"],
                       tree["block_groups"], None, np.zeros(3), theta, 0.1, csr=csr, keep_H=False)
    bounds, npar = synth.default_bounds(q), theta.size
    kw = dict(burn=0, thin=1, adapting=True, rng_mode=1, sample_predicts=False, save_w=True, save_yhat=False)
    gm.mcmc(bounds, np.eye(npar) * sd, keep=5, seed=4, **kw)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        r = gm.mcmc(bounds, np.eye(npar) * sd, keep=iters, seed=5, **kw)
        torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"timeline_{name}.json")
    prof.export_chrome_trace(path)
    report(path, name, f"{iters} iterations, accepted {r['n_accepted']}, mcmc_time {r['mcmc_time'] * 1e3:.3f} ms -> {r['mcmc_time'] * 1e6 / iters:.1f} us / iteration")
    gm.close()


if __name__ == "__main__":
    main()
