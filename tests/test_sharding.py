"""CPU tests of the multi-GPU host logic: block ranges never split an 8-block group, and scatter -> encode -> gather
over torch.distributed (gloo, world_size 2 and 3) reproduces the single-process result.  The per-rank encoder used
here is the CPU oracle standing in for the CUDA library (the collectives and range logic are what is under test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from convectionkernels_b200 import sharding, synth


def test_shard_ranges_cover_and_respect_groups():
    for n_groups in (1, 2, 7, 8, 9, 131072, 524288, 13):
        for world in (1, 2, 3, 4, 8):
            r = sharding.shard_ranges(n_groups * 8, world)
            assert len(r) == world
            pos = 0
            for first, n in r:
                assert first == pos and first % 8 == 0 and n % 8 == 0
                pos += n
            assert pos == n_groups * 8
            sizes = [n for _, n in r]
            assert max(sizes) - min(sizes) <= 8
    with pytest.raises(ValueError):
        sharding.shard_ranges(12, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_blocks, result_path):
    import sys
    sys.path.insert(0, ROOT)
    from oracle.loader import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O = Oracle()
        opt = np.zeros(44, np.uint8)
        plan = O.plan_from_quality(3)
        import struct
        opt[:] = np.frombuffer(struct.pack("<I5f5i", 0x108, 0.5, 0.2125 / 0.7154, 1.0, 0.0721 / 0.7154, 1.0, 2, 3, 8, 2, 4), np.uint8)
        blocks = torch.from_numpy(synth.random_blocks_rgba8(n_blocks, seed=77).reshape(-1)) if rank == 0 else None

        def enc(local):
            out = O.encode_bc7(local.numpy(), opt, plan)
            return torch.from_numpy(out.reshape(-1))

        full = sharding.encode_sharded(enc, blocks, n_blocks, 64, 16, src=0)
        if rank == 0:
            want = O.encode_bc7(blocks.numpy(), opt, plan)
            np.save(result_path, np.array([int((full.numpy() == want).all()), full.shape[0]]))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_blocks", [(2, 72), (3, 40)])
def test_scatter_encode_gather_gloo(tmp_path, world, n_blocks):
    path = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_blocks, path), nprocs=world, join=True)
    ok, n = np.load(path)
    assert ok == 1 and n == n_blocks
