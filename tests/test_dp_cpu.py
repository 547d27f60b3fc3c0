"""Host-side logic of the data-parallel step on CPU: world_size-2 gloo run of the replica exchange
(gradient sum + overflow-counter sum before the range controller) and the oracle-level
statement of why the counters must be summed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dfxp as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from lbt_b200.trainer import sync_replicas
    rng = np.random.default_rng(rank)
    flat_g = torch.from_numpy(rng.standard_normal(1000).astype(np.float32))
    # per-quantiser counters [n, 4] = (over, over_half, numel, ticket) measured on this replica's shard
    x = (rng.standard_normal((64, 32)) * (3.0 if rank == 1 else 0.4)).astype(np.float32)
    n1, n2 = O.overflow_counts(x, 8, 2)
    counters = torch.tensor([[n1, n2, x.size, 0], [0, 0, 0, 0]], dtype=torch.int64)
    sync_replicas(flat_g, counters, None, True)
    if rank == 0:
        torch.save(dict(g=flat_g, c=counters), out)
    dist.barrier()
    dist.destroy_process_group()


def test_sync_replicas_gloo_world2(tmp_path):
    port = _free_port()
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    gs, shards = [], []
    for r in range(2):
        rng = np.random.default_rng(r)
        gs.append(torch.from_numpy(rng.standard_normal(1000).astype(np.float32)))
        shards.append((rng.standard_normal((64, 32)) * (3.0 if r == 1 else 0.4)).astype(np.float32))
    assert torch.equal(got['g'], gs[0] + gs[1])
    # summed counters == counters of the global batch, so the controller decides as one device at batch N*b would
    full = np.concatenate(shards)
    n1, n2 = O.overflow_counts(full, 8, 2)
    assert got['c'][0].tolist() == [n1, n2, full.size, 0]
    assert got['c'][1].tolist() == [0, 0, 0, 0]
    want = O.Range(2)
    O.update_range(full, 0.0, 8, want)
    d = O.range_delta(int(got['c'][0, 0]), int(got['c'][0, 1]), int(got['c'][0, 2]), 0.0)
    assert min(7, 2 + d) == want.value
    # without the exchange replica 0 alone would have shrunk its range: the replicas would drift apart
    alone = O.Range(2)
    O.update_range(shards[0], 0.0, 8, alone)
    assert alone.value != want.value


def test_finalize_and_prep_job_structs_match_header():
    """ctypes mirrors of lbt_finalize_job / lbt_prep_job have the C layout (no padding surprises)."""
    import ctypes
    from lbt_b200 import _lib
    assert ctypes.sizeof(_lib.FinalizeJob) == 64
    assert ctypes.sizeof(_lib.PrepJob) == 128
    assert _lib.FinalizeJob.start.offset == 56 and _lib.FinalizeJob.exp_const.offset == 32
    assert _lib.PrepJob.bits.offset == 88 and _lib.PrepJob.rot180.offset == 116 and _lib.PrepJob.sw.offset == 124
