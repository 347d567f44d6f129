"""CPU, world_size 2 over gloo: the multi-GPU host logic (image sharding + the one result gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ctpn_b200.dist import gather_results, shard_range


def test_shard_range_covers_batch_contiguously():
    for n in (0, 1, 7, 32, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, post = 3, 5
    rois = torch.full((B, post, 5), float(rank), dtype=torch.float32)
    rois[:, :, 0] += torch.arange(B, dtype=torch.float32)[:, None] / 10
    count = torch.tensor([rank * 10 + b for b in range(B)], dtype=torch.int32)
    all_r, all_c = gather_results(rois, count)
    q.put((rank, all_r.numpy().copy(), all_c.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_results_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, all_r, all_c in got:
        assert all_r.shape == (6, 5, 5)
        np.testing.assert_array_equal(all_c, [0, 1, 2, 10, 11, 12])          # rank order, image order within rank
        np.testing.assert_allclose(all_r[:, 0, 1], [0, 0, 0, 1, 1, 1])
        np.testing.assert_allclose(all_r[:, 0, 0], [0.0, 0.1, 0.2, 1.0, 1.1, 1.2], rtol=1e-6)
