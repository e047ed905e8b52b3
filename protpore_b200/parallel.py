"""Multi-GPU: one process per GPU (torchrun), events sharded, NCCL only where the path exchanges data.

Scoring / decoding / posteriors need no communication: every event is independent and the model
(33 KB of CSR + emission parameters) is replicated.  `shard_events` deals events to ranks by length so
that every rank holds the same number of DP cells; results are per-event and independent of the world size.

The one collective of the path is Baum-Welch: after the per-GPU E-step
(`Model.baum_welch_statistics`, kernel `k_fb_accumulate`) the expected transition counts per edge, the
per-state emission moments (sum w, sum w x, sum w x^2) and the log-likelihood are SUM-allreduced as ONE
packed float64 buffer (1,151 doubles for pro_001), the observed ranges of Uniform states MAX-allreduced as
(-min, max); every rank then runs the identical M-step (`Model.apply_m_step`, yahmm.pyx:4253-4327).
"""
import math

import numpy as np


def shard_events(lengths, rank, world):
    """Indices of the events rank `rank` owns: longest first, dealt in a snake so cells balance."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    pos = np.arange(len(order))
    lap, slot = pos // world, pos % world
    owner = np.where(lap % 2 == 0, slot, world - 1 - slot)
    return np.sort(order[owner == rank])


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def _device_for_collectives(device=None):
    import torch
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device() if device is None else device)
    return torch.device("cpu")


def allreduce_statistics(edge_counts, moments, minmax, extra=None, device=None):
    """In-place allreduce of Baum-Welch sufficient statistics (numpy float64 arrays) over the default group.

    `extra` (optional 1-D float64 array, e.g. [sum log-likelihood, #events, #skipped]) is summed as well.
    No process group -> no-op (single GPU)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return
    import torch
    dev = _device_for_collectives(device)
    parts = [edge_counts.ravel(), moments.ravel()] + ([extra.ravel()] if extra is not None else [])
    packed = torch.from_numpy(np.concatenate(parts)).to(dev)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    rng = torch.from_numpy(np.concatenate([-minmax[:, 0], minmax[:, 1]])).to(dev)
    dist.all_reduce(rng, op=dist.ReduceOp.MAX)
    out = packed.cpu().numpy()
    n_e, n_m = edge_counts.size, moments.size
    edge_counts.ravel()[:] = out[:n_e]
    moments.ravel()[:] = out[n_e:n_e + n_m]
    if extra is not None:
        extra.ravel()[:] = out[n_e + n_m:]
    r = rng.cpu().numpy()
    ne = minmax.shape[0]
    minmax[:, 0] = -r[:ne]
    minmax[:, 1] = r[ne:]


def allreduce_logsumexp(local_values, device=None):
    """log sum exp over the values of ALL ranks (the reference's training score, yahmm.pyx:93-99)."""
    local_values = np.asarray(local_values, dtype=np.float64)
    finite = local_values[np.isfinite(local_values)]
    mx = float(finite.max()) if finite.size else -math.inf
    dist = _dist()
    if dist is not None and dist.get_world_size() > 1:
        import torch
        dev = _device_for_collectives(device)
        t = torch.tensor([mx], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mx = float(t.item())
    s = float(np.exp(finite - mx).sum()) if (finite.size and math.isfinite(mx)) else 0.0
    if dist is not None and dist.get_world_size() > 1:
        import torch
        t = torch.tensor([s], dtype=torch.float64, device=_device_for_collectives(device))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        s = float(t.item())
    return mx + math.log(s) if s > 0 else -math.inf


def baum_welch_iteration(model, local_sequences, transition_pseudocount=0, use_pseudocount=False, edge_inertia=0.0,
                         distribution_inertia=0.0, dtype=np.float32):
    """One distributed Baum-Welch iteration: local GPU E-step, allreduce, identical M-step on every rank.

    Returns (sum of log-likelihoods over all ranks' possible events, number of events used)."""
    logp, edge, mom, mm = model.baum_welch_statistics(local_sequences, dtype=dtype)
    ok = np.isfinite(logp)
    extra = np.array([float(logp[ok].sum()), float(ok.sum())], dtype=np.float64)
    allreduce_statistics(edge, mom, mm, extra)
    model.apply_m_step(edge, mom, mm, transition_pseudocount, int(use_pseudocount), edge_inertia, distribution_inertia)
    return float(extra[0]), int(extra[1])


def baum_welch_iteration_packed(model, obs, offsets, transition_pseudocount=0, use_pseudocount=False, edge_inertia=0.0,
                                distribution_inertia=0.0):
    """`baum_welch_iteration` on events already packed CSR-style (`obs` back to back, `offsets[n + 1]`): numpy arrays
    or torch CUDA tensors resident on the model's GPU.  One batched E-step call, one packed allreduce, the M-step."""
    logp, edge, mom, mm, _ = model.engine().forward_backward_batch(obs, offsets)
    if not isinstance(logp, np.ndarray):                      # device tensors: reduce the per-event values there
        import torch
        ok = torch.isfinite(logp)
        extra = np.array([float(logp[ok].sum()), float(ok.sum())], dtype=np.float64)
        edge, mom, mm = edge.cpu().numpy(), mom.cpu().numpy(), mm.cpu().numpy()
    else:
        ok = np.isfinite(logp)
        extra = np.array([float(logp[ok].sum()), float(ok.sum())], dtype=np.float64)
    allreduce_statistics(edge, mom, mm, extra)
    model.apply_m_step(edge, mom, mm, transition_pseudocount, int(use_pseudocount), edge_inertia, distribution_inertia)
    return float(extra[0]), int(extra[1])


def train_baum_welch(model, local_sequences, stop_threshold=1e-9, min_iterations=0, max_iterations=None, verbose=False,
                     **kwargs):
    """Distributed `Model.train(..., algorithm='baum-welch')`: same stopping rule as the reference
    (yahmm.pyx:4031-4073) on the log-sum-exp of ALL ranks' sequence log-probabilities."""
    last = allreduce_logsumexp(model.log_probability_batch(local_sequences))
    first, iteration, improvement = last, 0, float("inf")
    while improvement > stop_threshold or iteration < min_iterations:
        if max_iterations and iteration >= max_iterations:
            break
        baum_welch_iteration(model, local_sequences, **kwargs)
        iteration += 1
        now = allreduce_logsumexp(model.log_probability_batch(local_sequences))
        improvement, last = now - last, now
        if verbose:
            print("Training improvement: {}".format(improvement))
    return last - first
