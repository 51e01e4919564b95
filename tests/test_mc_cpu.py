"""
Host logic of the Monte-Carlo path that needs no GPU: packing the
per-instance device records through the reference's sweep mechanism,
sharding of instance ranges over ranks, and the end-of-run gather
(world_size 2, gloo).
"""
import os

import numpy as np
import pytest

from tests import support as S
from cardoon_b200.analyses import mc


def test_pack_instances_and_restore():
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = queue[0]
    bt = mc.BatchTransient(ckt, an.tstop, an.tstep, an.im)
    elems = ckt.nD_nlinElem
    before = [e.cuda_spec()['params'].copy() for e in elems]
    P = bt.pack_instances(5, mc.bsim3_ring_variation())
    assert len(P) == len(bt.topo.groups)
    ng = bt.topo.groups[0]['n']
    assert P[0].shape[1] == 5 * ng
    # a later range reproduces the same instance (seeded by instance number)
    P2 = bt.pack_instances(2, mc.bsim3_ring_variation(), first=3)
    assert np.array_equal(P2[0][:, :ng], P[0][:, 3 * ng:4 * ng])
    assert not np.array_equal(P[0][:, :ng], P[0][:, ng:2 * ng])
    # the circuit's own elements are untouched afterwards
    for e, b in zip(elems, before):
        assert np.array_equal(e.cuda_spec()['params'], b)
    # nominal "variation" reproduces the circuit's own records
    P0 = bt.pack_instances(1, lambda j, el: None)
    order = [e for g in bt.topo.groups for e in g['elems']]
    assert np.array_equal(P0[0][:, 0], order[0].cuda_spec()['params'])


def test_shard_ranges():
    for total, world in ((4096, 8), (4096, 3), (7, 2), (5, 8)):
        got = []
        for r in range(world):
            lo, hi = mc.shard(total, r, world)
            got += list(range(lo, hi))
        assert got == list(range(total))


def _gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    B = 3
    wave = torch.full((B, 4, 2), float(rank), dtype=torch.float64) \
        + torch.arange(B, dtype=torch.float64)[:, None, None] * 0.1
    iters = torch.full((B, 4), rank, dtype=torch.int32)
    status = torch.full((B,), rank, dtype=torch.int32)
    w, i, s = mc.gather_results(wave, iters, status, dist)
    q.put((rank, w.numpy(), i.numpy(), s.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_results_gloo_world2():
    import torch.multiprocessing as tmp
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q))
             for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, w, i, s in res:
        assert w.shape == (6, 4, 2) and i.shape == (6, 4)
        assert list(s) == [0, 0, 0, 1, 1, 1]
        assert np.allclose(w[:3, 0, 0], [0., .1, .2])
        assert np.allclose(w[3:, 0, 0], [1., 1.1, 1.2])
