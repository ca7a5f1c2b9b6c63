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
    # B200CV_AUTO_DP loss reporting: value = sum over the replicas (train.py:70 `losses[0].sum()` on a DataParallel
    # output), gradient = this replica's own
    w = torch.tensor([float(rank + 1)], requires_grad=True)
    tot = parallel.sum_over_replicas(w * 2.0)
    tot.sum().backward()
    torch.save({"arena": arena, "shard": (lo, hi), "dp_value": tot.detach(), "dp_grad": w.grad},
               os.path.join(out_dir, f"r{rank}.pt"))
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
    for r in res:
        assert float(r["dp_value"]) == 6.0 and float(r["dp_grad"]) == 2.0  # 2*1 + 2*2 everywhere; own gradient only


def test_single_process_is_a_noop():
    import sys

    from b200cv import parallel

    t = torch.ones(4)
    parallel.allreduce_gradients(t)
    assert torch.equal(t, torch.ones(4)) and parallel.world_size() == 1 and parallel.shard_batch(5) == (0, 5)


def test_dataparallel_chunks_and_gpu_binding(monkeypatch):
    """B200CV_AUTO_DP: the scatter chunks of nn.DataParallel (torch.chunk: ceil(n/w) per replica) and the one-GPU-per-
    process binding that lets the unchanged train.py (device 'cuda:0', DataParallel only when device_count() > 1) run
    under torchrun."""
    from b200cv import parallel

    assert [parallel.dp_chunk(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]
    assert [parallel.dp_chunk(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]  # torch.chunk sizes
    assert parallel.dp_chunk(5, 3, 4) == (5, 5)  # trailing replica without work
    monkeypatch.delenv("B200CV_AUTO_DP", raising=False)
    monkeypatch.setenv("WORLD_SIZE", "4")
    assert not parallel.auto_dp_enabled()
    monkeypatch.setenv("B200CV_AUTO_DP", "1")
    assert parallel.auto_dp_enabled()
    monkeypatch.setenv("LOCAL_RANK", "2")
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "4,5,6,7")
    monkeypatch.delenv("B200CV_AUTO_DP_BOUND", raising=False)
    monkeypatch.setattr(torch.cuda, "is_initialized", lambda: False)
    parallel.bind_process_to_local_gpu()
    assert os.environ["CUDA_VISIBLE_DEVICES"] == "6" and os.environ["B200CV_AUTO_DP_BOUND"] == "1"
    parallel.bind_process_to_local_gpu()  # DataLoader workers inherit the environment: no second remap
    assert os.environ["CUDA_VISIBLE_DEVICES"] == "6"
