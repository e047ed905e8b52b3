"""Multi-GPU: one process per GPU (torchrun), events sharded, NCCL only where the path exchanges data.

Scoring / decoding / posteriors need no communication: every event is independent and the model
(33 KB of CSR + emission parameters) is replicated.  `shard_events` deals events to ranks by length so
that every rank holds the same number of DP cells; results are per-event and independent of the world size.

The one collective of the path is Baum-Welch: after the per-GPU E-step (`k_fb_lane` /
`k_fb_accumulate`) the expected transition counts per edge, the per-state emission moments (sum w, sum w x,
sum w x^2), the observed ranges of the Uniform states and the log-likelihood travel as ONE packed float64 buffer
(1,409 doubles for pro_001) in ONE all_gather; every rank reduces the shards in rank order on its GPU (sums,
minima, maxima) and runs the identical M-step (`Model.apply_m_step`, yahmm.pyx:4253-4327).  On the device-resident
path the E-step kernels accumulate straight into that buffer (`PackedStatistics`): nothing is staged through the host
before the exchange, and one device -> host copy feeds the M-step.
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


def _gather(packed):
    """ONE collective: every rank's packed statistics, rank-major (all_gather_into_tensor).  Gathering instead of
    all-reducing lets the sums, minima and maxima share one exchange and makes the result independent of the
    collective's reduction order: every rank adds the shards in rank order and ends up bit-identical."""
    import torch
    dist = _dist()
    world = dist.get_world_size()
    out = torch.empty(world * packed.numel(), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.contiguous().view(-1))
    return out.view(world, packed.numel())


def allreduce_statistics(edge_counts, moments, minmax, extra=None, device=None):
    """In-place allreduce of Baum-Welch sufficient statistics (numpy float64 arrays) over the default group: edge
    counts, emission moments and `extra` (e.g. [sum log-likelihood, #events]) are summed, the observed ranges of the
    Uniform states reduced with min / max -- in ONE collective.  No process group -> no-op (single GPU)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return
    import torch
    dev = _device_for_collectives(device)
    parts = [edge_counts.ravel(), moments.ravel(), minmax[:, 0], minmax[:, 1]] + ([extra.ravel()] if extra is not None else [])
    g = _gather(torch.from_numpy(np.concatenate(parts)).to(dev))
    n_e, n_m, ne = edge_counts.size, moments.size, minmax.shape[0]
    total = g.sum(dim=0).cpu().numpy()
    lo = g[:, n_e + n_m:n_e + n_m + ne].amin(dim=0).cpu().numpy()
    hi = g[:, n_e + n_m + ne:n_e + n_m + 2 * ne].amax(dim=0).cpu().numpy()
    edge_counts.ravel()[:] = total[:n_e]
    moments.ravel()[:] = total[n_e:n_e + n_m]
    minmax[:, 0], minmax[:, 1] = lo, hi
    if extra is not None:
        extra.ravel()[:] = total[n_e + n_m + 2 * ne:]


class PackedStatistics(object):
    """The Baum-Welch statistics of one rank as ONE persistent device buffer: the E-step kernels accumulate straight into
    views of it (pp_forward_backward_batch takes caller-owned accumulators), so the buffer the collective sends is the
    buffer the reduction kernel wrote -- no staging copy, no host round trip before the exchange.

    Layout (float64): [edge counts E | moments 3 ne | range minima ne | range maxima ne | sum log-likelihood, #events]."""

    def __init__(self, n_edges, ne, device):
        import torch
        self.n_edges, self.ne = n_edges, ne
        self.buf = torch.zeros(n_edges + 5 * ne + 2, dtype=torch.float64, device=device)
        self._mm = torch.empty((ne, 2), dtype=torch.float64, device=device)     # the ABI's (min, max) pairs
        self.reset()

    def reset(self):
        self.buf.zero_()
        self._mm[:, 0] = float("inf")
        self._mm[:, 1] = float("-inf")

    @property
    def edge(self):
        return self.buf[:self.n_edges]

    @property
    def moments(self):
        return self.buf[self.n_edges:self.n_edges + 3 * self.ne].view(self.ne, 3)

    def accumulators(self):
        return self.edge, self.moments, self._mm

    def finish(self, logp):
        """Per-event log-likelihoods -> the two scalars; ranges into their columns.  All on the device."""
        import torch
        ok = torch.isfinite(logp)
        o = self.n_edges + 3 * self.ne
        self.buf[o:o + self.ne] = self._mm[:, 0]
        self.buf[o + self.ne:o + 2 * self.ne] = self._mm[:, 1]
        self.buf[-2] = torch.where(ok, logp, torch.zeros_like(logp)).sum()
        self.buf[-1] = ok.sum().to(torch.float64)

    def reduce(self):
        """The collective (one all_gather, skipped without a process group) and the only device -> host copy of the
        iteration.  Returns numpy (edge, moments, minmax, [sum log-likelihood, #events])."""
        dist = _dist()
        o = self.n_edges + 3 * self.ne
        if dist is not None and dist.get_world_size() > 1:
            g = _gather(self.buf)
            total = g.sum(dim=0)
            total[o:o + self.ne] = g[:, o:o + self.ne].amin(dim=0)
            total[o + self.ne:o + 2 * self.ne] = g[:, o + self.ne:o + 2 * self.ne].amax(dim=0)
        else:
            total = self.buf
        h = total.cpu().numpy()
        mm = np.stack([h[o:o + self.ne], h[o + self.ne:o + 2 * self.ne]], axis=1)
        return h[:self.n_edges].copy(), h[self.n_edges:o].reshape(self.ne, 3).copy(), mm, h[-2:].copy()


def allreduce_logsumexp(local_values, device=None):
    """log sum exp over the values of ALL ranks (the reference's training score, yahmm.pyx:93-99) in ONE collective:
    every rank contributes (its maximum, its sum of exponentials relative to that maximum)."""
    local_values = np.asarray(local_values, dtype=np.float64)
    finite = local_values[np.isfinite(local_values)]
    mx = float(finite.max()) if finite.size else -math.inf
    s = float(np.exp(finite - mx).sum()) if finite.size else 0.0
    dist = _dist()
    if dist is not None and dist.get_world_size() > 1:
        import torch
        g = _gather(torch.tensor([mx, s], dtype=torch.float64, device=_device_for_collectives(device))).cpu().numpy()
        top = float(g[:, 0].max())
        if not math.isfinite(top):
            return -math.inf
        s = float(sum(si * math.exp(mi - top) for mi, si in g if math.isfinite(mi)))
        mx = top
    return mx + math.log(s) if s > 0 else -math.inf


def _refuse_unreduced_families(model):
    """Emission families without sufficient statistics (kernel densities, mixtures, LogNormal ...) are re-estimated from
    the items and posterior weights of the sequences a rank holds (Distribution.summarize, yahmm.pyx:4215-4236); nothing
    here ships those across ranks, so training them on several GPUs would silently use one shard.  Refuse instead."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return
    for state in model.states[:model.silent_start]:
        d = state.distribution
        if not (hasattr(d, "summarize_moments") or hasattr(d, "summarize_range")) and not d.frozen:
            raise NotImplementedError(
                "distributed Baum-Welch re-estimates Normal and Uniform emissions; state {!r} has a trainable {} -- freeze it "
                "or train on one GPU".format(state.name, type(d).__name__))


def baum_welch_iteration(model, local_sequences, transition_pseudocount=0, use_pseudocount=False, edge_inertia=0.0,
                         distribution_inertia=0.0, dtype=np.float32):
    """One distributed Baum-Welch iteration: local GPU E-step, one collective, identical M-step on every rank.

    Returns (sum of log-likelihoods over all ranks' possible events, number of events used)."""
    _refuse_unreduced_families(model)
    logp, edge, mom, mm = model.baum_welch_statistics(local_sequences, dtype=dtype)
    ok = np.isfinite(logp)
    extra = np.array([float(logp[ok].sum()), float(ok.sum())], dtype=np.float64)
    allreduce_statistics(edge, mom, mm, extra)
    model.apply_m_step(edge, mom, mm, transition_pseudocount, int(use_pseudocount), edge_inertia, distribution_inertia)
    return float(extra[0]), int(extra[1])


def baum_welch_iteration_packed(model, obs, offsets, transition_pseudocount=0, use_pseudocount=False, edge_inertia=0.0,
                                distribution_inertia=0.0, stats=None):
    """`baum_welch_iteration` on events already packed CSR-style (`obs` back to back, `offsets[n + 1]`): numpy arrays
    or torch CUDA tensors resident on the model's GPU.  Device tensors take the device-side path: the E-step accumulates
    into one persistent packed buffer (`stats`, a PackedStatistics; made on first use), ONE collective gathers it, the
    reduction stays on the device and a single copy brings the result to the host M-step."""
    _refuse_unreduced_families(model)
    eng = model.engine()
    if isinstance(obs, np.ndarray):
        logp, edge, mom, mm, _ = eng.forward_backward_batch(obs, offsets)
        ok = np.isfinite(logp)
        extra = np.array([float(logp[ok].sum()), float(ok.sum())], dtype=np.float64)
        allreduce_statistics(edge, mom, mm, extra)
    else:
        if stats is None:
            stats = getattr(model, "_packed_statistics", None)
            if stats is None or stats.n_edges != eng.n_edges or stats.ne != eng.ne or stats.buf.device != obs.device:
                stats = PackedStatistics(eng.n_edges, eng.ne, obs.device)
                model._packed_statistics = stats
        stats.reset()
        logp = eng.forward_backward_batch(obs, offsets, accum=stats.accumulators())[0]
        stats.finish(logp)
        edge, mom, mm, extra = stats.reduce()
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
