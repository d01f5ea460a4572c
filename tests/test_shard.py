"""CPU: the N>1 host logic -- static sharding of independent units, world_size-2 gloo rendezvous, result gather.
There is no collective on the data path; gloo is only used here to run two real ranks."""
import os
import socket

import numpy as np
import pytest


def test_shard_units_partition():
    from open_speech_b200.batch import shard_units

    for n in (0, 1, 7, 256, 1024, 4097):
        for world in (1, 2, 4, 8):
            parts = [list(shard_units(n, world, r)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_length_balanced():
    from open_speech_b200.batch import shard_by_length

    rng = np.random.default_rng(1005)
    lens = rng.integers(48000, 288000, 4096)
    for world in (2, 4, 8):
        bins = shard_by_length(lens, world)
        assert sorted(sum(bins, [])) == list(range(4096))
        loads = [int(lens[b].sum()) for b in bins]
        assert (max(loads) - min(loads)) / max(loads) < 0.01


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, n_units, out_dir):
    import torch.distributed as dist

    from open_speech_b200.batch import shard_units

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(shard_units(n_units, world, rank))
    # stand-in for the per-rank hot path: every unit yields a value that depends only on the unit id
    local = [(u, (u * 2654435761) % 1000003) for u in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    dist.barrier()
    if rank == 0:
        flat = sorted(sum(gathered, []))
        np.save(os.path.join(out_dir, "gathered.npy"), np.array(flat))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    n_units, world = 257, 2
    mp.spawn(_rank_main, args=(world, _free_port(), n_units, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    assert got[:, 0].tolist() == list(range(n_units))
    assert got[:, 1].tolist() == [(u * 2654435761) % 1000003 for u in range(n_units)]
