"""The N>1 launch path on CPU: world_size-2 gloo, file list sharded by rank, digests gathered
on rank 0 by index.  The hasher here is the oracle (there is no GPU in this container); on
the GPU box the same sharding code feeds libsnapgpu (bench.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import oracle as O
    from snappy_b200 import sharding, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = synth.lognormal_sizes(600)
    shards = sharding.contiguous_shards(lengths, world)
    b, e = shards[rank]
    data, off, ln = synth.make_host_batch(lengths[b:e], first_index=b)       # this rank's files only
    local = O.sha512_batch(data, off, ln)
    full = sharding.gather_digests(local, shards, rank, world)
    if rank == 0:
        np.save(out_path, full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path, oracle):
    import torch.multiprocessing as mp
    from snappy_b200 import synth
    out = str(tmp_path / "digests.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    lengths = synth.lognormal_sizes(600)
    data, off, ln = synth.make_host_batch(lengths)
    assert np.array_equal(got, oracle.sha512_batch(data, off, ln, 4))


def test_contiguous_shards_are_balanced():
    from snappy_b200 import sharding, synth
    for lengths in (np.full(2_000_000, 65536, dtype=np.uint64), synth.lognormal_sizes(100_000),
                    np.array([5], dtype=np.uint64), np.zeros(0, dtype=np.uint64)):
        for world in (1, 2, 4, 8):
            sh = sharding.contiguous_shards(lengths, world)
            assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == len(lengths)
            assert all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))
            if len(lengths) >= 1000:
                blocks = synth.blocks(lengths).astype(np.float64)
                loads = [blocks[b:e].sum() for b, e in sh]
                assert max(loads) <= blocks.sum() / world * 1.001 + blocks.max()
