"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous sharding, one all-gather of
the 48-byte records, input order restored, results identical to the single-rank run."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sblx import shard


def _fake_results(lo, hi):
    idx = np.arange(lo, hi)
    return dict(converged=(idx % 7 != 0).astype(np.uint8), status=(idx % 7 == 0).astype(np.int32), iterations=(3 + idx % 5).astype(np.int32),
                n_switch_events=(idx % 3).astype(np.int32), final_mismatch=1e-9 * (1 + idx), vmin_pu=0.95 + 1e-4 * idx,
                vmax_pu=1.05 + 1e-4 * idx, max_loading_pct=50.0 + idx)


def _worker(rank, world, port, total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(total, rank, world)
    rec = shard.pack_records(_fake_results(lo, hi))
    t = torch.from_numpy(np.frombuffer(rec.tobytes(), dtype=np.uint8).copy())
    out = shard.gather_records(t, total, world, dist)
    got = shard.records_from_bytes(out)
    ref = shard.pack_records(_fake_results(0, total))
    ok = got.tobytes() == ref.tobytes()
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges_cover_in_order():
    for total in (0, 1, 7, 29601, 100000):
        for world in (1, 2, 4, 8):
            r = [shard.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(hi - lo for lo, hi in r) <= -(-total // world)


def test_two_rank_gather_restores_input_order():
    world, total = 2, 1001       # odd total: the last rank is padded
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, ret), nprocs=world, join=True)
    assert all(ret[k] for k in range(world))
