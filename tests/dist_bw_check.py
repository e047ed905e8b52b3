"""Worker of tests/test_multi_gpu.py: one Baum-Welch iteration on N GPUs (torchrun, NCCL) against one GPU and against the
compiled reference's golden iteration.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from protpore_b200 import parallel, events  # noqa: E402
import protpore_b200.yahmm as y  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    y.Model.device = local
    out = {}
    # ---- (1) the golden Baum-Welch iteration of the compiled reference, events sharded over the ranks
    g = util.load_golden("pro_001_py2")
    evs = util.golden_events(g)
    seqs = [evs[i] for i in g["bw_events"]]
    model = util.build_model("pro_001_py2")
    mine = parallel.shard_events([len(s) for s in seqs], rank, world)
    parallel.baum_welch_iteration(model, [seqs[i] for i in mine], dtype=np.float64)
    dense = model.dense_transition_matrix()
    fin = np.isfinite(g["bw_dense"])
    params = np.array([[float(p) for p in s.distribution.parameters[:2]] for s in model.states[:model.silent_start]])
    out["golden_dense_ok"] = bool(np.array_equal(np.isfinite(dense), fin) and np.allclose(dense[fin], g["bw_dense"][fin], rtol=1e-8, atol=1e-9))
    out["golden_params_ok"] = bool(np.allclose(params, g["bw_params"], rtol=1e-7, atol=1e-9))
    # ---- (2) device-resident path: statistics of N ranks' shards (one all_gather) vs all events on one GPU
    model = util.build_model("pro_001_py2")
    eng = model.engine()
    n = 6000
    lens = events.uniform_lengths(n, 50, 300, seed=5)
    obs, off = eng.sample_events(lens, seed=6)
    mine = parallel.shard_events(lens, rank, world)
    sub = [obs[off[i]:off[i + 1]] for i in mine]
    from protpore_b200 import _engine
    o_l, off_l = _engine.pack_events(sub, np.float32)
    dev = torch.device("cuda", local)
    stats = parallel.PackedStatistics(eng.n_edges, eng.ne, dev)
    logp = eng.forward_backward_batch(torch.from_numpy(o_l).to(dev), torch.from_numpy(off_l).to(dev), accum=stats.accumulators())[0]
    stats.finish(logp)
    edge, mom, mm, extra = stats.reduce()
    lp1, e1, m1, mm1, _ = eng.forward_backward_batch(obs, off)          # all events, this GPU alone
    S = float(n)
    tol = 1e-12 * np.sqrt(S)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)[...]) if a.size else 0.0)
    big = np.abs(e1) > 1e-6 * np.abs(e1).max()
    out["edge_rel"] = rel(edge[big], e1[big])
    bigm = np.abs(m1) > 1e-6 * np.abs(m1).max()
    out["moment_rel"] = rel(mom[bigm], m1[bigm])
    out["minmax_equal"] = bool(np.array_equal(mm, mm1))
    out["loglik_rel"] = abs(extra[0] - lp1[np.isfinite(lp1)].sum()) / abs(extra[0])
    out["events"] = int(extra[1])
    out["tol"] = tol
    # every rank holds the identical reduced buffer
    t = torch.from_numpy(np.concatenate([edge, mom.ravel()])).to(dev)
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([1.0 if torch.equal(t, ref) else 0.0], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out["ranks_identical"] = bool(same.item() == 1.0)
    out["world"] = world
    if rank == 0:
        print("DIST_BW " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
