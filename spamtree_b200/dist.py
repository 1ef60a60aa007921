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


def partitioned_model(d, tree, theta, beta, tausq, rank, nranks, device, allreduce, keep_H=False, native_nccl=True, limited_tree=False):
    """SpamTreeMV of this rank's share of the problem (d: data dict with y/X/coords/mv_id/q; tree: make_tree output).
    limited_tree: every block conditions on its direct parent only (make_edges_limited); the subtrees are still cut along
    the full ancestor chains."""
    pl = part.plan(tree, d["y"], nranks)
    mp = None
    if limited_tree:
        from .api import limited_edges_csr
        lp = limited_edges_csr(tree, d["y"])
        mp = (lp[0], lp[1])
    sp = part.subproblem(d, tree, pl, rank, nranks, model_parents=mp)
    sp["allreduce"] = allreduce
    gm = SpamTreeMV(sp["y"], sp["X"], sp["coords"], sp["mv_id"], sp["res_is_ref"], None, None, bool(limited_tree), sp["block_names"],
                    sp["block_groups"], None, beta, theta, tausq, csr=sp["csr"], device=device, keep_H=keep_H,
                    partition=sp if nranks > 1 else None, q=d["q"])
    if nranks > 1 and native_nccl:
        attach_native_nccl(gm, rank)
    return gm, sp, pl


def nccl_unique_id():
    import ctypes as C
    from ._lib import lib
    buf = C.create_string_buffer(128)
    rc = lib.st_nccl_unique_id(buf)
    if rc:
        raise RuntimeError("st_nccl_unique_id failed: " + lib.st_last_error(None).decode())
    return buf.raw


def attach_native_nccl(gm, rank):
    """gives the partitioned handle its own NCCL communicator: rank 0 draws the unique id, torch.distributed hands it to
    the other ranks, every rank attaches (collective).  Only with the nccl backend (a gloo group has no GPUs to talk to)."""
    import torch
    import torch.distributed as dist
    if dist.get_backend() != "nccl":
        return False
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    gm.attach_nccl(box[0])
    return True
