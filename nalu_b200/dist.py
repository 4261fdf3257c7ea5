"""torch.distributed plumbing for the multi-rank path: one process per GPU.

The only exchanges on the path are (1) the once-per-mesh graph handshake, host side
(TpetraLinearSystem.C:678-738,924-952 does it with stk::CommNeighbors) and (2) loadComplete's
shared-row export-add (:1757-1778), device side: packed doubles over NCCL send/recv (NVLink), or
read straight out of the neighbour's arrays through CUDA-IPC mapped peer pointers (halo_pull_add).
`gloo` on CPU carries (1) in the test-suite; (2) needs devices."""
import numpy as np


def make_exchange(rank, world, group=None):
    """Neighbour handshake over torch.distributed: send {dest: int64 array} -> {src: int64 array}."""
    import torch.distributed as dist

    def exchange(send):
        allsend = [None] * world
        dist.all_gather_object(allsend, {int(k): np.asarray(v, dtype=np.int64) for k, v in send.items()}, group=group)
        return {src: allsend[src][rank] for src in range(world) if allsend[src] and rank in allsend[src]}

    return exchange


def exchange_halo_plan(ls):
    """(send, recv) peers of a finalized linear system as [(rank, n_doubles)] -- what the NCCL leg
    of loadComplete posts.  Pure host logic, usable without a device."""
    return ls.halo_peers()


class PeerWindows:
    """CUDA-IPC mapped views of every neighbour's shared-not-owned block, for the pull variant of
    loadComplete: the owner's add kernel reads the sharer's CSR values over NVLink, no pack, no
    staging buffer, no unpack."""

    def __init__(self, ls, rank, world, group=None):
        import ctypes as C

        import torch.distributed as dist

        from .lib import check, lib
        L = lib()
        self.ls = ls
        self.L = L
        h = (C.c_ubyte * 64)()
        check(L.nalu_b200_ipc_get_handle(C.c_void_p(ls.values.data_ptr()), h), "ipc_get_handle")
        hv = bytes(h)
        check(L.nalu_b200_ipc_get_handle(C.c_void_p(ls.rhs.data_ptr()), h), "ipc_get_handle")
        hr = bytes(h)
        # the linear system must have been allocated with allocate(raw=True): a handle names a whole
        # cudaMalloc block, which is the tensor itself only when it did not come from a caching allocator
        if not getattr(ls, "_raw", None):
            raise RuntimeError("PeerWindows needs DeviceCsrLinearSystem.allocate(raw=True)")
        mine = dict(values=hv, rhs=hr, nnz_owned=int(ls.nnz_owned), n_owned_rows=int(ls.n_owned_rows), voff=0, roff=0)
        allh = [None] * world
        dist.all_gather_object(allh, mine, group=group)
        self.ptrs = {}
        self.opened = []
        _, recv = ls.halo_peers()
        for i, (peer, _) in enumerate(recv):
            pv, pr = C.c_void_p(), C.c_void_p()
            check(L.nalu_b200_ipc_open_handle(allh[peer]["values"], C.byref(pv)), "ipc_open_handle")
            check(L.nalu_b200_ipc_open_handle(allh[peer]["rhs"], C.byref(pr)), "ipc_open_handle")
            self.opened += [pv, pr]
            self.ptrs[i] = (pv.value + allh[peer]["voff"] + 8 * allh[peer]["nnz_owned"],
                            pr.value + allh[peer]["roff"] + 8 * allh[peer]["n_owned_rows"])

    def load_complete(self, barrier):
        """`barrier()` = synchronise this rank's stream, then barrier across ranks.  Every rank's assembly must be
        finished before its rows are pulled, and no rank may start its next assembly (overwrite mode rewrites the
        shared rows) or zeroSystem while an owner is still reading them over NVLink: barrier on both sides."""
        barrier()
        for i, (pv, pr) in self.ptrs.items():
            self.ls.halo_pull_add(i, pv, pr)
        barrier()

    def close(self):
        for p in self.opened:
            self.L.nalu_b200_ipc_close(p)
        self.opened = []




def make_comm(rank, world):
    """The library's own NCCL communicator (nalu_b200_comm_create): rank 0's unique id travels over torch.distributed."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import lib as _lib
    L = _lib.lib()
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_char * 128)()
        _lib.check(L.nalu_b200_comm_get_unique_id(buf), "comm_get_unique_id")
        ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    dev = torch.device("cuda", torch.cuda.current_device())
    t = ident.to(dev)
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().numpy().tobytes())
    comm = C.c_void_p()
    _lib.check(L.nalu_b200_comm_create(C.byref(comm), world, rank, raw), "comm_create")
    return comm
