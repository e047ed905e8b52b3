"""Host-side multi-GPU logic on CPU: event sharding, the packed allreduce of Baum-Welch statistics and the
distributed training score, with world_size 2 over gloo (what the NCCL path does on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from protpore_b200 import parallel


def test_shard_events_partitions_and_balances():
    lens = np.random.RandomState(0).randint(20, 20000, size=1001)
    for world in (1, 2, 4, 8):
        shards = [parallel.shard_events(lens, r, world) for r in range(world)]
        allidx = np.sort(np.concatenate(shards))
        assert np.array_equal(allidx, np.arange(len(lens)))              # a partition
        cells = np.array([lens[s].sum() for s in shards], dtype=np.float64)
        assert cells.max() / cells.min() < 1.02                          # balanced to 2 %
        assert all(np.all(np.diff(s) > 0) for s in shards)                # original order kept inside a shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = util.build_model("synth_12")
        ne, E = model.silent_start, model.edge_count()
        rng = np.random.RandomState(100 + rank)
        edge, mom = rng.uniform(0, 5, size=E), rng.uniform(0, 3, size=(ne, 3))
        mom[:, 1] = mom[:, 0] * rng.uniform(25, 50, size=ne)
        mom[:, 2] = mom[:, 1] ** 2 / np.maximum(mom[:, 0], 1e-9) + mom[:, 0] * rng.uniform(0.2, 2.0, size=ne)
        mm = np.stack([rng.uniform(20, 30, size=ne), rng.uniform(40, 60, size=ne)], axis=1)
        if rank == 1:
            mm[3] = [np.inf, -np.inf]                                   # a state this rank never saw
        extra = np.array([-100.0 * (rank + 1), 10.0 + rank])
        local = (edge.copy(), mom.copy(), mm.copy(), extra.copy())
        parallel.allreduce_statistics(edge, mom, mm, extra)
        lse = parallel.allreduce_logsumexp(np.array([-5.0 - rank, -7.0, -np.inf]))
        model.apply_m_step(edge, mom, mm)
        # a trainable family without sufficient statistics cannot be re-estimated from one rank's shard: refused
        import protpore_b200.yahmm as y
        kd = y.Model(name="kd")
        st = y.State(y.GaussianKernelDensity([1.0, 2.0, 3.0], 1.0), name="k")
        kd.add_state(st)
        kd.add_transition(kd.start, st, 1.0); kd.add_transition(st, st, 0.5); kd.add_transition(st, kd.end, 0.5)
        kd.bake()
        refused = False
        try:
            parallel._refuse_unreduced_families(kd)
        except NotImplementedError:
            refused = True
        assert refused
        parallel._refuse_unreduced_families(model)                      # Normal / Uniform only: fine
        params = np.array([[float(p) for p in s.distribution.parameters[:2]] for s in model.states[:ne]])
        np.savez(os.path.join(tmp, "rank%d.npz" % rank), edge=edge, mom=mom, mm=mm, extra=extra, lse=lse,
                 dense=model.dense_transition_matrix(), params=params, l_edge=local[0], l_mom=local[1], l_mm=local[2],
                 l_extra=local[3])
    finally:
        dist.destroy_process_group()


def test_allreduce_statistics_and_identical_m_step_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("edge", "mom", "mm", "extra", "dense", "params"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k             # every rank ends up identical
    assert np.allclose(a["edge"], a["l_edge"] + b["l_edge"])
    assert np.allclose(a["mom"], a["l_mom"] + b["l_mom"])
    assert np.allclose(a["extra"], a["l_extra"] + b["l_extra"])
    assert np.array_equal(a["mm"][:, 0], np.minimum(a["l_mm"][:, 0], b["l_mm"][:, 0]))
    assert np.array_equal(a["mm"][:, 1], np.maximum(a["l_mm"][:, 1], b["l_mm"][:, 1]))
    want = np.log(np.exp(-5.0) + np.exp(-7.0) + np.exp(-6.0) + np.exp(-7.0))
    assert float(a["lse"]) == pytest.approx(want, rel=1e-12) and float(b["lse"]) == pytest.approx(want, rel=1e-12)


def test_single_process_is_a_noop():
    edge, mom = np.ones(5), np.ones((3, 3))
    mm = np.array([[1.0, 2.0]] * 3)
    parallel.allreduce_statistics(edge, mom, mm)
    assert edge.sum() == 5 and mom.sum() == 9
    assert parallel.allreduce_logsumexp([-1.0, -1.0]) == pytest.approx(-1.0 + np.log(2.0))
