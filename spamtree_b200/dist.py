"""torch.distributed plumbing of the multi-GPU partition: the library's allreduce callback and a one-call constructor."""
import numpy as np

from . import partition as part
from .api import SpamTreeMV


class _DevBuf:
    """lets torch alias `count` doubles at a raw device pointer (CUDA array interface)"""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def make_allreduce(device=None):
    """sum over the ranks of the default process group, in place, of `count` doubles at device pointer `ptr`.
    NCCL: the buffer is aliased as a CUDA tensor; gloo: it is staged through the host."""
    import torch
    import torch.distributed as dist

    def allreduce(ptr, count):
        t = torch.as_tensor(_DevBuf(ptr, count), device=device)
        if dist.get_backend() == "nccl":
            dist.all_reduce(t)
            torch.cuda.current_stream(t.device).synchronize()
        else:
            h = t.cpu()
            dist.all_reduce(h)
            t.copy_(h)
            torch.cuda.synchronize(t.device)

    return allreduce


def partitioned_model(d, tree, theta, beta, tausq, rank, nranks, device, allreduce, keep_H=False):
    """SpamTreeMV of this rank's share of the problem (d: data dict with y/X/coords/mv_id/q; tree: make_tree output)"""
    pl = part.plan(tree, d["y"], nranks)
    sp = part.subproblem(d, tree, pl, rank, nranks)
    sp["allreduce"] = allreduce
    gm = SpamTreeMV(sp["y"], sp["X"], sp["coords"], sp["mv_id"], sp["res_is_ref"], None, None, False, sp["block_names"],
                    sp["block_groups"], None, beta, theta, tausq, csr=sp["csr"], device=device, keep_H=keep_H,
                    partition=sp if nranks > 1 else None, q=d["q"])
    return gm, sp, pl
