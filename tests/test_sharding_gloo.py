"""N>1 host logic on CPU: two gloo ranks shard one batch by contiguous packet range (no data-path collective),
each decodes its shard, and the gathered result equals the single-process decode. The decode itself is done by
the oracle here (no GPU in this suite); what is under test is the product's partitioning and ordering."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from signals import make_signal


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _workload():
    cfg = ol.Config.make(bit_depth=16, num_channels=2, sample_rate=44100, frame_length=1024)
    x = make_signal('silence_lsb', 2, 1024 * 37 + 100, 16, 44100, seed=11)
    return cfg, ol.encode_stream(cfg, x)


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from alac_b200_loader import load_package
    pkg = load_package()
    cfg, packets = _workload()
    lo, hi = pkg.shard_ranges([len(p) for p in packets], world)[rank]
    hashes = []
    for p in packets[lo:hi]:
        st, pcm = ol.decode_packet(cfg, p)
        assert st == 0
        hashes.append(hashlib.sha256(pcm).hexdigest())
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, hashes))  # control plane only: ranges + digests
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, packets = _workload()
    want = [hashlib.sha256(ol.decode_packet(cfg, p)[1]).hexdigest() for p in packets]
    gathered.sort()
    assert gathered[0][0] == 0 and gathered[0][1] == gathered[1][0] and gathered[1][1] == len(packets)
    assert gathered[0][2] + gathered[1][2] == want
    sizes = np.array([len(p) for p in packets])
    a, b = sizes[:gathered[0][1]].sum(), sizes[gathered[0][1]:].sum()
    assert abs(int(a) - int(b)) <= 2 * sizes.max()  # balanced by compressed bytes


def test_shard_ranges_edge_cases():
    from alac_b200_loader import load_package
    pkg = load_package()
    assert pkg.shard_ranges([], 4) == [(0, 0)] * 4
    assert pkg.shard_ranges([5], 3) in ([(0, 0), (0, 0), (0, 1)], [(0, 1), (1, 1), (1, 1)], [(0, 0), (0, 1), (1, 1)])
    r = pkg.shard_ranges([1] * 10, 8)
    assert r[0][0] == 0 and r[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    r = pkg.shard_ranges([100, 1, 1, 1, 1, 100], 2)
    assert r == [(0, 3), (3, 6)] or r == [(0, 1), (1, 6)] or r == [(0, 2), (2, 6)]
