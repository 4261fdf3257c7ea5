"""N>1 host logic over a real process group (gloo, CPU): every rank builds only its own slab, the graph
handshake travels through torch.distributed exactly as bench.py / a user would run it under torchrun, and
each rank's graph, row/column maps and halo plan must equal the oracle's simulated MPI job."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dims, ndof, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch.distributed as dist

        import oracle
        from nalu_b200 import dist as nd
        from nalu_b200 import mesh
        from nalu_b200.linsys import DeviceCsrLinearSystem
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        m = mesh.generated_box_rank(*dims, rank, world)          # this rank's slab only
        ls = DeviceCsrLinearSystem(m, ndof, exchange=nd.make_exchange(rank, world), host_only=True)
        parts = [mesh.generated_box_rank(*dims, r, world) for r in range(world)]
        orc = oracle.LinSys(parts, ndof)                          # the checker simulates the whole job
        assert ls.n_owned_rows == orc.n_owned_rows(rank) and ls.n_shared_rows == orc.n_shared_rows(rank)
        for shared in (False, True):
            np.testing.assert_array_equal(ls.rowptr(shared), orc.rowptr(rank, shared))
            np.testing.assert_array_equal(ls.colidx(shared), orc.colidx(rank, shared))
            np.testing.assert_array_equal(ls.row_gids(shared), orc.row_gids(rank, shared))
        np.testing.assert_array_equal(ls.col_gids(), orc.col_gids(rank))
        send, recv = nd.exchange_halo_plan(ls)
        # slab decomposition: shared rows go to the lower neighbour only, and what I send is what it expects
        assert [p for p, _ in send] == ([rank - 1] if rank > 0 else [])
        assert [p for p, _ in recv] == ([rank + 1] if rank < world - 1 else [])
        plans = [None] * world
        dist.all_gather_object(plans, (send, recv))
        for p, n in send:
            assert (rank, n) in plans[p][1]
        if send:
            assert send[0][1] == ls.nnz_shared + ls.n_shared_rows
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc() + str(e)))


@pytest.mark.parametrize("world,dims,ndof", [(2, (3, 2, 4), 3), (3, (4, 3, 7), 1)])
def test_graph_handshake_over_gloo(world, dims, ndof):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, ndof, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in res), res
