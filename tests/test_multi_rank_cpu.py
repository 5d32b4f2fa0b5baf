"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups run the sharding + all-gather plumbing of
openmavis_b200/shard.py; the 2-NN compute callback is the oracle (the GPU callback is tested in test_gpu_match.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from openmavis_b200 import shard
from openmavis_b200.synth import synth_descriptors, synth_query_from_train


def test_frame_shard_partitions_exactly():
    for n in (0, 1, 7, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            blocks = [shard.frame_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        train = synth_descriptors(n, 5); query = synth_query_from_train(train, 6)
        lo, cnt = shard.row_shard(n, rank, world)
        t_local = torch.from_numpy(train[lo:lo + cnt].copy())
        idx, dd = shard.knn2_sharded(query[lo:lo + cnt], t_local, n, dist,
                                     lambda q, t: O.knn2(np.ascontiguousarray(q), t.numpy(), 2))
        full_idx, full_dist = shard.gather_results(idx, dd, n, dist)
        if rank == 0:
            np.savez(out, idx=full_idx, dist=full_dist)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n', [(2, 601), (3, 500)])
def test_sharded_knn2_equals_single_rank(tmp_path, world, n):
    out = str(tmp_path / 'res.npz')
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    got = np.load(out)
    train = synth_descriptors(n, 5); query = synth_query_from_train(train, 6)
    ridx, rdist = O.knn2(query, train)
    assert np.array_equal(got['idx'], ridx) and np.array_equal(got['dist'], rdist)
