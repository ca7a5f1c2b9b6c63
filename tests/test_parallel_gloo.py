"""Data-parallel plumbing on CPU: world_size-2 gloo processes exercise init_from_env, the SUM gradient
all-reduce (DataParallel semantics: sum of per-replica gradients, CVC-YOLOv3/train.py:70,193-195) and
batch sharding."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "mit-driverless-cv-traininginfra_b200"))
    from b200cv import parallel

    local = parallel.init_from_env(backend="gloo")
    assert local == rank and parallel.world_size() == world and parallel.rank() == rank
    arena = torch.arange(10, dtype=torch.float32) * (rank + 1)  # the flat gradient arena of this replica
    parallel.allreduce_gradients(arena)
    lo, hi = parallel.shard_batch(7)
    torch.save({"arena": arena, "shard": (lo, hi)}, os.path.join(out_dir, f"r{rank}.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_allreduce_sum_and_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    want = torch.arange(10, dtype=torch.float32) * 3  # SUM, not mean
    for r in res:
        assert torch.equal(r["arena"], want)
    assert [r["shard"] for r in res] == [(0, 4), (4, 7)]


def test_single_process_is_a_noop():
    import sys

    from b200cv import parallel

    t = torch.ones(4)
    parallel.allreduce_gradients(t)
    assert torch.equal(t, torch.ones(4)) and parallel.world_size() == 1 and parallel.shard_batch(5) == (0, 5)
