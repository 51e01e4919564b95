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


def _gather_worker(rank, world, port, q, sizes):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    B = sizes[rank]
    wave = torch.full((B, 4, 2), float(rank), dtype=torch.float64) \
        + torch.arange(B, dtype=torch.float64)[:, None, None] * 0.1
    iters = torch.full((B, 4), rank, dtype=torch.int32)
    status = torch.full((B,), rank, dtype=torch.int32)
    w, i, s = mc.gather_results(wave, iters, status, dist)
    q.put((rank, w.numpy(), i.numpy(), s.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _run_gather(sizes, port_base):
    import torch.multiprocessing as tmp
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    port = port_base + os.getpid() % 200
    procs = [ctx.Process(target=_gather_worker,
                         args=(r, len(sizes), port, q, sizes))
             for r in range(len(sizes))]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_gather_results_gloo_world2():
    for rank, w, i, s in _run_gather((3, 3), 29650):
        assert w.shape == (6, 4, 2) and i.shape == (6, 4)
        assert list(s) == [0, 0, 0, 1, 1, 1]
        assert np.allclose(w[:3, 0, 0], [0., .1, .2])
        assert np.allclose(w[3:, 0, 0], [1., 1.1, 1.2])


@pytest.mark.parametrize('sizes', [(4, 3), (0, 2)])
def test_gather_results_unequal_shards(sizes):
    """shard() gives unequal (even empty) ranges when total % world != 0."""
    for rank, w, i, s in _run_gather(sizes, 29860 + 7 * sizes[0]):
        tot = sum(sizes)
        assert w.shape == (tot, 4, 2) and i.shape == (tot, 4)
        assert list(s) == [0] * sizes[0] + [1] * sizes[1]
        assert np.allclose(w[sizes[0]:, 0, 0], 1. + .1 * np.arange(sizes[1]))


def test_pack_batch_equals_scalar_packer():
    """The vectorised packer is bit-identical to the per-instance reference
    mechanism (setattr + process_params + cuda_spec per element)."""
    ckt, queue = S.load_circuit(S.net('ring_bsim3.net'))
    an = queue[0]
    bt = mc.BatchTransient(ckt, an.tstop, an.tstep, an.im)
    elems = ckt.nD_nlinElem
    before = [e.cuda_spec()['params'].copy() for e in elems]
    ids = list(range(300)) + [31, 4095, 2048]
    z = mc.bsim3_ring_samples(ids)
    fast = bt.pack_batch(mc.bsim3_ring_overrides(z, elems))
    slow = bt.pack_instances(len(ids), mc.bsim3_ring_variation(),
                             instances=ids)
    for a, b in zip(fast, slow):
        assert a.shape == b.shape
        assert np.array_equal(a, b)
    for e, b in zip(elems, before):             # circuit left untouched
        assert np.array_equal(e.cuda_spec()['params'], b)
    # a variation that would change the kernel structure is refused
    bad = mc.bsim3_ring_overrides(z, elems)
    bad[0]['alpha0'] = np.where(np.arange(len(ids)) % 2, 1e-6, 0.)
    with pytest.raises(NotImplementedError):
        bt.pack_batch(bad)


GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'mc_ring_bsim3.npz')


def test_mc_golden_against_live_oracle():
    """tests/golden/mc_ring_bsim3.npz (make_mc_golden.py) is what the oracle
    computes today: operating-point helper / iterations of a few instances,
    among them 31 (non-finite Jacobian on the way, gmin stepping), and the
    first samples of one transient."""
    from tests.golden import make_mc_golden as G
    g = np.load(GOLD)
    assert g['op_helper'].shape == (4096,)
    assert g['op_helper'][31] == 1          # plain Newton fails, gmin2 works
    hard = np.nonzero(g['op_helper'] != 0)[0]
    assert 50 < len(hard) < 400 and set(hard[:32]) <= set(g['inst'])
    for j in (0, 31, int(hard[5])):
        _, k, it, x = G.op_of(j)
        assert k == g['op_helper'][j] and it == g['op_iters'][j]
        at = list(g['inst']).index(j)
        assert np.allclose(x, g['op_x'][at], rtol=1e-12, atol=1e-14)
    # transient head of instance 31
    from oracle import nodal_ref as R
    from cardoon_b200.globalVars import glVar
    ckt, an = G._instance(31)
    r = R.run_tran(ckt, glVar, an.tstop, an.tstep, an.im, max_samples=25)
    rows = [r['idx'].row(ckt.find_term(t)) for t in G.TERMS]
    at = list(g['inst']).index(31)
    assert np.array_equal(r['iters'], g['iters'][at][:25])
    assert np.allclose(r['X'][:, rows], g['wave'][at][:25], rtol=1e-12,
                       atol=1e-14)
