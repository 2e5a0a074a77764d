"""World-size-2 gloo test of the N>1 host logic: column sharding of one flightline, flightline sharding of
a batch, and the final gather.  The per-shard compute is the CPU oracle here (the CUDA path needs a GPU);
what is under test is that sharding + gather reproduce the unsharded result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from srcfinder_b200 import dist as cdist


def test_column_shard_partitions():
    for S in (598, 1242, 7, 8, 75):
        for world in (1, 2, 4, 8):
            parts = [cdist.column_shard(S, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == S
            for a, b in zip(parts[:-1], parts[1:]):
                assert a[1] == b[0]
            widths = [b - a for a, b in parts]
            assert max(widths) - min(widths) <= 1
    assert cdist.column_shard(598, 8, 0) == (0, 75) and cdist.column_shard(598, 8, 7) == (524, 598)
    assert cdist.flightline_shard(64, 8, 3) == list(range(3, 64, 8))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cmf_oracle as orc
    from srcfinder_b200 import synth
    active = [351, 422]
    ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
    cube = synth.make_cube(200, 7, seed=77, bad_pixels=True)          # every rank builds the same cube
    # (i) column sharding of one flightline
    s0, s1 = cdist.column_shard(cube.shape[2], world, rank)
    part = orc.cmf_cube(cdist.slice_columns(cube, s0, s1), ab, active)
    full = cdist.gather_column_tiles(torch.from_numpy(part["mf"]), cube.shape[2], dst=0)
    # (ii) flightline sharding of a batch of 3
    mine = cdist.flightline_shard(3, world, rank)
    tiles = [torch.from_numpy(orc.cmf_cube(synth.make_cube(120, 3, seed=90 + f), ab, active)["mf"]) for f in mine]
    batch = cdist.gather_flightlines(tiles, 3, mine, dst=0)
    # (iii) fewer flightlines than ranks: rank 1 owns none and still takes part in the collective
    lone = cdist.flightline_shard(1, world, rank)
    lone_tiles = [tiles[0]] if (rank == 0 and lone) else []
    single = cdist.gather_flightlines(lone_tiles, 1, lone, dst=0)
    if rank == 0:
        assert len(single) == 1 and torch.equal(single[0], tiles[0])
        assert lone == [0]
    else:
        assert lone == [] and single is None
    if rank == 0:
        ref = orc.cmf_cube(cube, ab, active)["mf"]
        np.save(os.path.join(tmp, "ok.npy"), np.array([
            float(np.array_equal(full.numpy(), ref, equal_nan=True)),
            float(all(np.array_equal(batch[f].numpy(),
                                     orc.cmf_cube(synth.make_cube(120, 3, seed=90 + f), ab, active)["mf"])
                      for f in range(3)))]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok[0] == 1.0, "column-sharded gather differs from the unsharded result"
    assert ok[1] == 1.0, "flightline-sharded gather differs"


def test_even_column_shards_cover_the_image():
    for S in (598, 1242, 7, 16):
        for world in (1, 2, 3, 4, 8):
            sh = [cdist.column_shard_even(S, world, r) for r in range(world)]
            assert sh[0][0] == 0 and sh[-1][1] == S
            assert all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))
            if S % 2 == 0:
                assert all(a % 2 == 0 for a, _ in sh)        # 8-byte aligned rows for every shard of a BIL cube
