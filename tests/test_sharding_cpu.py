"""CPU tests of the multi-GPU host logic: partition arithmetic and the rendezvous / gather plumbing of
codes_b200.run_sharded, exercised with two gloo ranks (no GPU)."""
import os
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import codes_b200
from tests import oracle_util

ROOT = oracle_util.ROOT


@pytest.mark.parametrize("G,a,p", [(9, 4, 2), (33, 8, 4), (65, 16, 8), (257, 32, 16), (513, 64, 32)])
def test_partitions_tile_the_network(engine_lib, G, a, p):
    for world in (1, 2, 3, 4, 8):
        if world > G:
            continue
        prev_n = prev_r = 0
        sizes = []
        for rank in range(world):
            n0, n1, r0, r1 = codes_b200.partition_bounds(G, a, p, rank, world)
            assert (n0, r0) == (prev_n, prev_r)          # contiguous, no gaps, no overlap
            assert (n1 - n0) == (r1 - r0) * p            # terminals follow their routers
            assert (r1 - r0) % a == 0                    # whole groups only: local links never cross a rank
            sizes.append((r1 - r0) // a)
            prev_n, prev_r = n1, r1
        assert (prev_n, prev_r) == (G * a * p, G * a)
        assert max(sizes) - min(sizes) <= 1              # imbalance of at most one group
    with pytest.raises(codes_b200.EngineError):
        codes_b200.partition_bounds(G, a, p, 0, G + 1)
    with pytest.raises(codes_b200.EngineError):
        codes_b200.partition_bounds(G, a, p, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the same collectives run_sharded uses: all-gather of the 64-byte IPC handles, gather of the partitions
        handle = bytes([rank]) * 64
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        bounds = codes_b200.partition_bounds(33, 8, 4, rank, world)
        parts = [None] * world
        dist.gather_object((rank, bounds), parts if rank == 0 else None, dst=0)
        dist.barrier()
        q.put((rank, handles, parts))
    finally:
        dist.destroy_process_group()


def test_rendezvous_plumbing_two_gloo_ranks(engine_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29533 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, handles, parts in res:
        assert handles == [bytes([k]) * 64 for k in range(world)]   # rank order preserved: base[k] is rank k's region
        if rank == 0:
            assert [p[0] for p in parts] == [0, 1]
            assert parts[0][1][1] == parts[1][1][0]                 # rank 1 starts where rank 0 ends
        else:
            assert parts == [None] * world


def test_partition_api_guards(engine_lib):
    with codes_b200.Engine() as e:
        with pytest.raises(codes_b200.EngineError):
            e.set_partition(3, 2)
        e.set_partition(1, 2)
        with pytest.raises(codes_b200.EngineError, match="engine_upload"):
            e.ipc_export()
        with pytest.raises(codes_b200.EngineError, match="engine_finish"):
            e.export_partition()


def test_ensemble_hint_and_local_peer_guards(engine_lib):
    """dfd_engine_set_ensemble: engines that share one GPU (also the ranks of a sharded run held by one process);
    attach_local_peers needs uploaded engines."""
    with codes_b200.Engine() as e:
        e.set_ensemble(8)
        e.set_ensemble(0)  # back to the default
        e.set_partition(1, 2)
        e.set_ensemble(2)
        import ctypes
        arr = (ctypes.c_void_p * 2)(e._h, e._h)
        assert e._lib.dfd_engine_attach_local_peers(e._h, arr, 2) != 0
        assert "engine_upload" in e._lib.dfd_last_error(e._h).decode()
