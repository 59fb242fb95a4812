"""CPU: the multi-GPU host logic (clip sharding + final best-init gather) with world_size 2 over gloo."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from homan_b200 import distributed as hd


def test_shard_clips_partitions_every_clip_exactly_once():
    for n in (1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            got = sum((hd.shard_clips(n, r, w) for r in range(w)), [])
            assert got == list(range(n))
            sizes = [len(hd.shard_clips(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, num_clips, inits, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    losses = torch.from_numpy(np.random.default_rng(0).uniform(size=(num_clips, inits)).astype(np.float32))
    mine = hd.shard_clips(num_clips, rank, world)
    bi, bl = hd.local_best(losses[mine].reshape(-1), len(mine), inits)
    payload = torch.stack([torch.full((4,), float(c)) for c in mine]) if mine else torch.zeros(0, 4)
    gi, gl, gp = hd.gather_best(mine, bi, bl, num_clips, payload)
    torch.save((gi, gl, gp), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gather_best_world_size_2(tmp_path):
    num_clips, inits, world = 5, 16, 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, num_clips, inits, str(tmp_path)), nprocs=world, join=True)
    losses = np.random.default_rng(0).uniform(size=(num_clips, inits)).astype(np.float32)
    for r in range(world):
        gi, gl, gp = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert np.array_equal(gi.numpy(), losses.argmin(1))
        assert np.allclose(gl.numpy(), losses.min(1))
        assert np.array_equal(gp[:, 0].numpy(), np.arange(num_clips, dtype=np.float32))


def test_single_process_path():
    losses = torch.tensor([[3.0, 1.0, 2.0], [0.5, 4.0, 0.25]])
    bi, bl = hd.local_best(losses.reshape(-1), 2, 3)
    gi, gl, _ = hd.gather_best([0, 1], bi, bl, 2)
    assert gi.tolist() == [1, 2] and gl.tolist() == [1.0, 0.25]
