"""`Model` of the drop-in yahmm API (yahmm.pyx:2021-4499), with the DP sweeps on the GPU.

Graph construction, `bake`, the Baum-Welch M-step and the text format are host Python mirroring the
reference; every dynamic-programming sweep (`forward`, `backward`, `forward_backward`,
`log_probability`, `viterbi`, `maximum_a_posteriori`, the E-step of `train`) is one call into the
C-ABI engine (protpore_b200/_engine.py -> libprotpore_b200.so, hand-written sm_100a kernels).
There is no CPU implementation of those sweeps in this package: without the library and a CUDA
device they raise.

Added batched entry points (one GPU call for many events):
  `viterbi_batch(sequences)`, `log_probability_batch(sequences)`, `forward_backward_batch(sequences)`,
  `baum_welch_statistics(sequences)`.
"""
from __future__ import print_function

import math
import random
from functools import reduce

import numpy

from . import _graph
from .bake import bake_graph
from .distributions import Distribution, NEGINF
from .state import State


def _pair_lse(x, y):
    """yahmm.pyx:52-70."""
    if x == float("inf") or y == float("inf"):
        return float("inf")
    if x == NEGINF:
        return y
    if y == NEGINF:
        return x
    if x > y:
        return x + math.log(math.exp(y - x) + 1)
    return y + math.log(math.exp(x - y) + 1)


def log(value):
    """Natural log, -inf at 0; scalars and numpy arrays (yahmm.pyx:73-84)."""
    if isinstance(value, numpy.ndarray):
        out = numpy.zeros(value.shape)
        out[value > 0] = numpy.log(value[value > 0])
        out[value == 0] = NEGINF
        return out
    return math.log(value) if value > 0 else NEGINF


def exp(value):
    """yahmm.pyx:86-91."""
    return numpy.exp(value)


def log_probability(model, samples):
    """Log-sum-exp of the per-sequence log probabilities (yahmm.pyx:93-99), one batched GPU call."""
    batch = getattr(model, "log_probability_batch", None)
    if batch is not None and len(samples) > 0 and not isinstance(samples[0], tuple):
        return reduce(_pair_lse, [float(v) for v in batch(samples, dtype=numpy.float64)])   # the reference's precision
    return reduce(_pair_lse, map(model.log_probability, samples))


class _PackedSequences(list):
    """A list of sequences that remembers its packed form (observations back to back, offsets) for `Model._pack`."""
    packed = None


class Model(object):
    """Hidden Markov Model: states in a graph, lowered by `bake`, swept on the GPU."""

    device = 0      # CUDA device the next engine is created on (class default; instances may override)

    def __init__(self, name=None, start=None, end=None):
        self.name = name or str(id(self))
        self.graph = _graph.OrderedDiGraph()
        self.start = start or State(None, name=self.name + "-start")
        self.end = end or State(None, name=self.name + "-end")
        self.graph.add_node(self.start)
        self.graph.add_node(self.end)
        self.states = []
        self.start_index = self.end_index = self.silent_start = -1
        self._baked = None
        self._engine = None

    def __str__(self):
        return "{}:\n\t{}".format(self.name, "\n\t".join(map(str, self.states)))

    # ---- introspection (yahmm.pyx:2076-2115) -------------------------------------------------
    def state_count(self):
        return len(self.states)

    def edge_count(self):
        return len(self._baked.out_logp)

    def dense_transition_matrix(self):
        b = self._baked
        m = len(self.states)
        dense = numpy.zeros((m, m)) + NEGINF
        for i in range(m):
            for n in range(b.out_ptr[i], b.out_ptr[i + 1]):
                dense[i, b.out_idx[n]] = b.out_logp[n]
        return dense

    def is_infinite(self):
        return self._baked.finite == 0

    # ---- graph construction (yahmm.pyx:2117-2255) ----------------------------------------------
    def add_state(self, state):
        self.graph.add_node(state)

    def add_states(self, states):
        for state in states:
            self.add_state(state)

    def add_transition(self, a, b, probability, pseudocount=None, group=None):
        pseudocount = pseudocount or probability
        self.graph.add_edge(a, b, weight=log(probability), pseudocount=pseudocount, group=group)

    def add_transitions(self, a, b, probabilities=None, pseudocounts=None, groups=None):
        pseudocounts = pseudocounts or probabilities
        n = len(a) if isinstance(a, list) else len(b)
        if groups is None or isinstance(groups, str):
            groups = [groups] * n
        if isinstance(a, list) and isinstance(b, list):
            for s, e, p, c, g in zip(a, b, probabilities, pseudocounts, groups):
                self.add_transition(s, e, p, c, g)
        elif isinstance(a, list) and isinstance(b, State):
            for s, p, c, g in zip(a, probabilities, pseudocounts, groups):
                self.add_transition(s, b, p, c, g)
        elif isinstance(a, State) and isinstance(b, list):
            for e, p, c, g in zip(b, probabilities, pseudocounts, groups):
                self.add_transition(a, e, p, c, g)

    def add_model(self, other):
        self.graph = _graph.union(self.graph, other.graph)

    def concatenate_model(self, other):
        self.graph = _graph.union(self.graph, other.graph)
        self.add_transition(self.end, other.start, 1.00)
        self.end = other.end

    def draw(self, **kwargs):
        raise NotImplementedError("Model.draw needs networkx + matplotlib; plotting is outside the GPU path")

    def freeze_distributions(self):
        for state in self.states:
            if state.distribution is not None:
                state.distribution.freeze()

    def thaw_distributions(self):
        for state in self.states:
            if state.distribution is not None:
                state.distribution.thaw()

    # ---- bake (yahmm.pyx:2280-2655) ------------------------------------------------------------
    def bake(self, verbose=False, merge="all"):
        """Finalize the topology: lowers the graph to CSR (yahmm/bake.py) and drops any GPU handle."""
        self._baked = bake_graph(self, verbose=verbose, merge=merge)
        self.states = self._baked.states
        self.silent_start = self._baked.silent_start
        self.start_index = self._baked.start_index
        self.end_index = self._baked.end_index
        self._drop_engine()

    # ---- engine plumbing -------------------------------------------------------------------------
    def _drop_engine(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def _device_params(self):
        trip = [s.distribution.device_params() for s in self.states[:self.silent_start]]
        return [t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip]

    def engine(self):
        """The GPU handle of this baked model (created on first use; fails loudly without CUDA)."""
        if self._baked is None:
            raise RuntimeError("Model must be baked before it can be used")
        if self._engine is None:
            from .._engine import Engine
            kinds, p0, p1 = self._device_params()
            self._engine = Engine(self._baked, kinds, p0, p1, device=self.device)
        return self._engine

    def _sweep_engine(self, obs):
        """The engine, with the emission table of this batch in place when the model has states whose family the
        kernels do not evaluate themselves (kind 2): column k = log_probability of state k's
        distribution for every observation, the matrix the reference fills at yahmm.pyx:2836-2840."""
        eng = self.engine()
        kinds = self._device_params()[0]
        table_states = [k for k, kind in enumerate(kinds) if kind == 2]
        if table_states:
            x = numpy.asarray(obs.cpu() if hasattr(obs, "cpu") else obs, dtype=numpy.float64)
            table = numpy.zeros((len(x), self.silent_start), dtype=numpy.float64)
            done = {}
            for k in table_states:
                dist = self.states[k].distribution
                if id(dist) not in done:                       # tied states share one distribution object
                    done[id(dist)] = dist.log_probability_many(x)
                table[:, k] = done[id(dist)]
            eng.set_emission_table(table)
        return eng

    def _push_params(self):
        if self._engine is not None:
            kinds, p0, p1 = self._device_params()
            self._engine.update_params(self._baked, kinds, p0, p1)

    @staticmethod
    def _seq(sequence):
        x = numpy.ascontiguousarray(numpy.array(sequence), dtype=numpy.float64)
        if x.ndim != 1 or x.shape[0] == 0:
            raise ValueError("Invalid shape in axis 0: 0.")        # the reference's cvarray error (yahmm.pyx:2836)
        return x

    @staticmethod
    def _pack(sequences, dtype):
        from .._engine import pack_events
        packed = getattr(sequences, "packed", None)             # _PackedSequences: packed once per training run
        if packed is not None and packed[0].dtype == numpy.dtype(dtype):
            return packed
        for s in sequences:
            if len(s) == 0:
                raise ValueError("Invalid shape in axis 0: 0.")
        return pack_events(sequences, dtype)

    # ---- single-sequence API (same signatures and return values as the reference) --------------
    def forward(self, sequence):
        """Full forward table, (n+1) x m log probabilities (yahmm.pyx:2780-2930)."""
        return self._sweep_engine(self._seq(sequence)).forward_table(self._seq(sequence))

    def backward(self, sequence):
        """Full backward table, (n+1) x m log probabilities (yahmm.pyx:2932-3156)."""
        return self._sweep_engine(self._seq(sequence)).backward_table(self._seq(sequence))

    def _logp_from_forward(self, f, n):
        if self._baked.finite == 1:
            return float(f[n, self.end_index])
        s = NEGINF
        for i in range(self.silent_start):
            s = _pair_lse(s, float(f[n, i]))
        return s

    def log_probability(self, sequence, path=None):
        """log P(sequence) (yahmm.pyx:3366-3396), or of a sequence along a given path (3398-3442)."""
        if path:
            return self._log_probability_of_path(numpy.array(sequence), path)
        x = self._seq(sequence)
        return self._logp_from_forward(self._sweep_engine(x).forward_table(x), len(x))

    def _log_probability_of_path(self, sequence, path):
        b = self._baked
        indices = {s: i for i, s in enumerate(self.states)}
        i = 0
        log_score = 0.
        for j in range(1, len(path)):
            ki, ji = indices[path[j - 1]], indices[path[j]]
            found = False
            for l in range(b.out_ptr[ki], b.out_ptr[ki + 1]):
                if b.out_idx[l] == ji:
                    log_score += b.out_logp[l]
                    found = True
                    break
            if not found:
                return NEGINF
            if not path[j].is_silent():
                log_score += path[j].distribution.log_probability(sequence[i])
                i += 1
        return log_score

    def viterbi(self, sequence):
        """(log probability of the ML path, [(state_index, State), ...]) or (-inf, None) (yahmm.pyx:3444-3652)."""
        x = self._seq(sequence)
        offsets = numpy.array([0, len(x)], dtype=numpy.int64)
        scores, plen, paths = self._sweep_engine(x).viterbi_batch(x, offsets)
        if plen[0] < 0:
            return (float(scores[0]), None)
        return (float(scores[0]), [(int(k), self.states[k]) for k in paths[:plen[0]]])

    def _dense_expected(self, edge_counts):
        b = self._baked
        m = len(self.states)
        dense = numpy.zeros((m, m))
        src = numpy.repeat(numpy.arange(m), numpy.diff(b.out_ptr))
        dense[src, b.out_idx] = edge_counts
        return dense

    def forward_backward(self, sequence, tie=False):
        """(expected transitions m x m, log emission weights n x silent_start) or (None, None) (yahmm.pyx:3158-3364)."""
        x = self._seq(sequence)
        offsets = numpy.array([0, len(x)], dtype=numpy.int64)
        logp, edge, _, _, post = self._sweep_engine(x).forward_backward_batch(x, offsets, posteriors=True)
        if logp[0] == NEGINF:
            print("Warning: Sequence is impossible.")
            return (None, None)
        if tie:
            self._tie_posteriors(post)
        return self._dense_expected(edge), post

    def _tie_posteriors(self, ew):
        """LSE-pool log posteriors across states sharing a distribution (yahmm.pyx:3325-3361)."""
        b = self._baked
        visited = numpy.zeros(self.silent_start, dtype=bool)
        for k in range(self.silent_start):
            if visited[k]:
                continue
            visited[k] = True
            mates = list(b.tied_idx[b.tied_ptr[k]:b.tied_ptr[k + 1]])
            if not mates:
                continue
            visited[mates] = True
            for i in range(ew.shape[0]):
                acc = ew[i, k]
                for li in mates:
                    acc = _pair_lse(acc, ew[i, li])
                ew[i, mates] = acc
                ew[i, k] = acc

    def maximum_a_posteriori(self, sequence):
        """Posterior decoding (yahmm.pyx:3654-3742)."""
        x = self._seq(sequence)
        offsets = numpy.array([0, len(x)], dtype=numpy.int64)
        logp, _, _, _, post = self._sweep_engine(x).forward_backward_batch(x, offsets, posteriors=True)
        if logp[0] == NEGINF:
            print("Warning: Sequence is impossible.")
            return (None, None)
        path = [(self.start_index, self.start)]
        total = 0.
        for k in range(len(x)):
            best, arg = NEGINF, -1
            for l in range(self.silent_start):
                if post[k, l] > best:
                    best, arg = post[k, l], l
            path.append((arg, self.states[arg]))
            total += best
        path.append((self.end_index, self.end))
        return total, path

    # ---- batched entry points (new) --------------------------------------------------------------
    def log_probability_batch(self, sequences, dtype=numpy.float32):
        """log P of many sequences in one GPU call (scaled linear-space sweep, exact re-run where needed)."""
        obs, offsets = self._pack(sequences, dtype)
        return self._sweep_engine(obs).forward_batch(obs, offsets)

    def viterbi_batch(self, sequences, dtype=numpy.float32, states=True):
        """Viterbi of many sequences in one GPU call.

        Returns a list of what `viterbi(seq)` returns per sequence; with `states=False` the path items
        are bare state indices (numpy uint16 arrays) instead of (index, State) tuples."""
        obs, offsets = self._pack(sequences, dtype)
        scores, plen, paths = self._sweep_engine(obs).viterbi_batch(obs, offsets)
        out, pos = [], 0
        for e in range(len(sequences)):
            if plen[e] < 0:
                out.append((float(scores[e]), None))
                continue
            p = paths[pos:pos + plen[e]]
            pos += int(plen[e])
            out.append((float(scores[e]), [(int(k), self.states[k]) for k in p] if states else p.copy()))
        return out

    def forward_backward_batch(self, sequences, dtype=numpy.float32):
        """`forward_backward(seq)` for many sequences: list of (expected transitions, emission weights)."""
        out = []
        for s in sequences:          # per-event dense outputs; statistics for training use baum_welch_statistics
            obs, offsets = self._pack([s], dtype)
            logp, edge, _, _, post = self._sweep_engine(obs).forward_backward_batch(obs, offsets, posteriors=True)
            out.append((None, None) if logp[0] == NEGINF else (self._dense_expected(edge), post))
        return out

    def baum_welch_statistics(self, sequences, dtype=numpy.float32, accum=None):
        """E-step over many sequences in one GPU call: (logp[n], edge_counts[E], moments[ne,3], minmax[ne,2])."""
        obs, offsets = self._pack(sequences, dtype)
        logp, edge, mom, mm, _ = self._sweep_engine(obs).forward_backward_batch(obs, offsets, accum=accum)
        return logp, edge, mom, mm

    # ---- sampling (yahmm.pyx:2656-2778) ----------------------------------------------------------
    def sample(self, length=0, path=False):
        """Ancestral sample.  With `length` on a finite model, transitions into the end state are masked
        (and the rest renormalised) until `length` symbols are out -- what the reference intends at
        2742-2771 (its own loop reuses a stale index there)."""
        b = self._baked
        probs = numpy.exp(b.out_logp)
        i = self.start_index
        emissions, states = [], []
        while i != self.end_index:
            state = self.states[i]
            states.append(state)
            if not state.is_silent():
                emissions.append(state.distribution.sample())
                if length != 0 and len(emissions) >= length:
                    return [emissions, states] if path else emissions
            lo, hi = b.out_ptr[i], b.out_ptr[i + 1]
            w = probs[lo:hi].copy()
            dst = b.out_idx[lo:hi]
            if length != 0 and b.finite == 1 and not (hi - lo == 1 and dst[0] == self.end_index):
                w[dst == self.end_index] = 0.
            if w.sum() <= 0:
                break
            u = random.random() * w.sum()
            i = int(dst[min(int(numpy.searchsorted(numpy.cumsum(w), u, side='right')), hi - lo - 1)])
        if path:
            states.append(self.end)
            return [emissions, states]
        return emissions

    # ---- training (yahmm.pyx:3938-4499) ----------------------------------------------------------
    def train(self, sequences, stop_threshold=1E-9, min_iterations=0, max_iterations=None, algorithm='baum-welch',
              verbose=True, transition_pseudocount=0, use_pseudocount=False, edge_inertia=0.0,
              distribution_inertia=0.0):
        """Re-estimate the parameters from sequences; returns the total improvement (yahmm.pyx:3938-4029)."""
        use_pseudocount = int(use_pseudocount)
        algo = algorithm.lower()
        if algo in ('labelled', 'labeled'):
            sequences = [(numpy.array(s[0]), s[1]) for s in sequences]
            before = reduce(_pair_lse, [self.log_probability(s, p) for s, p in sequences])
            self._train_labelled(sequences, transition_pseudocount, use_pseudocount, edge_inertia,
                                 distribution_inertia)
            after = reduce(_pair_lse, [self.log_probability(s, p) for s, p in sequences])
        else:
            sequences = [numpy.asarray(s, dtype=numpy.float64) for s in sequences]
            before = log_probability(self, sequences)
            if algo == 'viterbi':
                self._train_viterbi(sequences, transition_pseudocount, use_pseudocount, edge_inertia,
                                    distribution_inertia)
            elif algo == 'baum-welch':
                self._train_baum_welch(sequences, stop_threshold, min_iterations, max_iterations, verbose,
                                       transition_pseudocount, use_pseudocount, edge_inertia, distribution_inertia)
            after = log_probability(self, sequences)
        improvement = after - before
        if verbose:
            print("Total Training Improvement: ", improvement)
        return improvement

    def _train_baum_welch(self, sequences, stop_threshold, min_iterations, max_iterations, verbose,
                          transition_pseudocount, use_pseudocount, edge_inertia, distribution_inertia):
        iteration, improvement = 0, float("+inf")
        # every iteration sweeps the same observations twice (E-step, then log_probability): copy them back to back ONCE
        sequences = _PackedSequences(sequences)
        sequences.packed = self._pack(list(sequences), numpy.float64)
        last = log_probability(self, sequences)
        while improvement > stop_threshold or iteration < min_iterations:
            if max_iterations and iteration >= max_iterations:
                break
            self._train_once_baum_welch(sequences, transition_pseudocount, use_pseudocount, edge_inertia,
                                        distribution_inertia)
            iteration += 1
            trained = log_probability(self, sequences)
            improvement = trained - last
            last = trained
            if verbose:
                print("Training improvement: {}".format(improvement))

    def _train_once_baum_welch(self, sequences, transition_pseudocount, use_pseudocount, edge_inertia,
                               distribution_inertia, allreduce=None):
        """One Baum-Welch iteration (yahmm.pyx:4075-4327): GPU E-step, optional allreduce, host M-step.

        `allreduce(edge, moments, minmax)` (in place) is where protpore_b200.parallel sums the per-GPU
        statistics over NCCL before every rank runs the identical M-step."""
        kinds = self._device_params()[0]
        table_states = [k for k, kind in enumerate(kinds) if kind == 2 and not self.states[k].distribution.frozen]
        if not table_states:
            _, edge, mom, mm = self.baum_welch_statistics(sequences, dtype=numpy.float64)
        else:
            # families without sufficient statistics keep the reference's route (4215-4236): the posterior weights of
            # every observation go to distribution.summarize(items, weights); tied states pool their weights
            obs, offsets = self._pack(sequences, numpy.float64)
            logp, edge, mom, mm, post = self._sweep_engine(obs).forward_backward_batch(obs, offsets, posteriors=True)
            used = numpy.repeat(numpy.isfinite(logp), numpy.diff(offsets))       # impossible sequences are skipped (4131-4133)
            b = self._baked
            seen = set()
            for k in table_states:
                dist = self.states[k].distribution
                if id(dist) in seen:
                    continue
                seen.add(id(dist))
                group = [k] + [int(j) for j in b.tied_idx[b.tied_ptr[k]:b.tied_ptr[k + 1]]]
                w = numpy.exp(post[:, group][used]).sum(axis=1)
                dist.summarize(obs[used], w)
        if allreduce is not None:
            allreduce(edge, mom, mm)
        self.apply_m_step(edge, mom, mm, transition_pseudocount, use_pseudocount, edge_inertia, distribution_inertia)

    def apply_m_step(self, edge_counts, moments, minmax, transition_pseudocount=0, use_pseudocount=0,
                     edge_inertia=0.0, distribution_inertia=0.0):
        """M-step from E-step sufficient statistics (yahmm.pyx:4253-4327, 489-523, 328-348)."""
        b = self._baked
        m = len(self.states)
        # plain Python lists: the arithmetic below is the reference's, operation by operation and in its order (the sums
        # that normalise a state's out-edges are sequential), and numpy scalar indexing would cost more than the arithmetic
        out_ptr, out_idx, in_ptr, in_idx = b.out_ptr.tolist(), b.out_idx.tolist(), b.in_ptr.tolist(), b.in_idx.tolist()
        out_pseudo, in_pseudo = b.out_pseudo.tolist(), b.in_pseudo.tolist()
        out_logp, in_logp = b.out_logp.tolist(), b.in_logp.tolist()
        cache = self.__dict__.get("_mstep_edge_of")            # (source, destination) -> out-edge index, per bake
        if cache is None or cache[0] is not b:
            edge_of = {}
            for k in range(m):
                for l in range(out_ptr[k], out_ptr[k + 1]):
                    edge_of[(k, out_idx[l])] = l
            self.__dict__["_mstep_edge_of"] = (b, edge_of)
        else:
            edge_of = cache[1]
        expected = [float(v) for v in edge_counts]
        # tied edge groups pool their counts (4253-4266)
        sizes = b.tied_edge_group_size
        for g in range(len(sizes) - 1):
            members = [edge_of[(int(b.tied_edges_starts[l]), int(b.tied_edges_ends[l]))] for l in range(sizes[g], sizes[g + 1])]
            total = sum(expected[e] for e in members)
            for e in members:
                expected[e] = total
        norm = [0.0] * m
        for k in range(m):
            acc = 0.0
            for l in range(out_ptr[k], out_ptr[k + 1]):
                acc += expected[l] + transition_pseudocount + out_pseudo[l] * use_pseudocount
            norm[k] = acc
        keep = 1 - edge_inertia
        for k in range(m):
            if norm[k] > 0:
                for l in range(out_ptr[k], out_ptr[k + 1]):
                    p = (expected[l] + transition_pseudocount + out_pseudo[l] * use_pseudocount) / norm[k]
                    out_logp[l] = log(math.exp(out_logp[l]) * edge_inertia + p * keep)
            for l in range(in_ptr[k], in_ptr[k + 1]):
                li = in_idx[l]
                if norm[li] > 0:
                    p = (expected[edge_of[(li, k)]] + transition_pseudocount + in_pseudo[l] * use_pseudocount) / norm[li]
                    in_logp[l] = log(math.exp(in_logp[l]) * edge_inertia + p * keep)
        b.out_logp[:] = out_logp
        b.in_logp[:] = in_logp
        # emission distributions, once per tied group, members' statistics pooled (4215-4236, 4304-4327)
        visited = numpy.zeros(self.silent_start, dtype=bool)
        for k in range(self.silent_start):
            if visited[k]:
                continue
            visited[k] = True
            dist = self.states[k].distribution
            if b.tied_ptr[k] == b.tied_ptr[k + 1]:             # not tied (the common case): no pooling, no numpy reductions
                sw, swx, swxx = moments[k]
                lo, hi = minmax[k]
            else:
                group = [k] + [int(j) for j in b.tied_idx[b.tied_ptr[k]:b.tied_ptr[k + 1]]]
                visited[group] = True
                sw, swx, swxx = moments[group, 0].sum(), moments[group, 1].sum(), moments[group, 2].sum()
                lo, hi = minmax[group, 0].min(), minmax[group, 1].max()
            if hasattr(dist, "summarize_moments"):
                dist.summarize_moments(sw, swx, swxx)
                dist.from_summaries(inertia=distribution_inertia)
            elif hasattr(dist, "summarize_range"):
                dist.summarize_range(sw, lo, hi)
                dist.from_summaries(inertia=distribution_inertia)
            elif not dist.frozen:
                # a table-state family: _train_once_baum_welch already handed it items and posterior weights
                # (Distribution.summarize); frozen ones (LambdaDistribution by default) are left alone as in the reference
                if dist.summaries:
                    # yahmm.pyx:4326 passes inertia=distribution_inertia to every family; the ones that inherit the base
                    # from_summaries (the kernel densities, 220) do not take it -- there the reference raises TypeError
                    try:
                        dist.from_summaries(inertia=distribution_inertia)
                    except TypeError:
                        dist.from_summaries()
        self._push_params()

    def _train_viterbi(self, sequences, transition_pseudocount, use_pseudocount, edge_inertia, distribution_inertia):
        """Hard-assignment training (yahmm.pyx:4329-4357): tag every sequence with ONE batched Viterbi call, then
        train on the (sequence, path) pairs as labelled data."""
        pairs = []
        for seq, (score, path) in zip(sequences, self.viterbi_batch(sequences, dtype=numpy.float64, states=False)):
            if path is None:
                print("Warning: skipped impossible sequence {}".format(seq))
                continue
            pairs.append((seq, [self.states[k] for k in path]))
        self._train_labelled(pairs, transition_pseudocount, use_pseudocount, edge_inertia, distribution_inertia)

    def _train_labelled(self, sequences, transition_pseudocount, use_pseudocount, edge_inertia, distribution_inertia):
        """Training on (sequence, state path) pairs (yahmm.pyx:4359-4499)."""
        b = self._baked
        m = len(self.states)
        indices = {s: i for i, s in enumerate(self.states)}
        counts = numpy.zeros((m, m), dtype=numpy.int64)
        symbols = [[] for _ in range(m)]
        for seq, labels in sequences:
            counts[self.start_index, indices[labels[0]]] += 1
            for x, y in zip(labels[:-1], labels[1:]):
                counts[indices[x], indices[y]] += 1
            counts[indices[labels[-1]], self.end_index] += 1
            i = 0
            for label in labels:
                if label.is_silent():
                    continue
                k = indices[label]
                symbols[k].append(seq[i])
                for li in b.tied_idx[b.tied_ptr[k]:b.tied_ptr[k + 1]]:
                    symbols[int(li)].append(seq[i])
                i += 1
        sizes = b.tied_edge_group_size                      # tied edges pool their counts (4425-4440)
        for g in range(len(sizes) - 1):
            members = [(int(b.tied_edges_starts[l]), int(b.tied_edges_ends[l])) for l in range(sizes[g], sizes[g + 1])]
            total = sum(counts[k] for k in members)
            for k in members:
                counts[k] = total
        norm = numpy.zeros(m)
        for k in range(m):
            for l in range(b.out_ptr[k], b.out_ptr[k + 1]):
                norm[k] += counts[k, b.out_idx[l]] + transition_pseudocount + b.out_pseudo[l] * use_pseudocount
        for k in range(m):
            if norm[k] > 0:
                for l in range(b.out_ptr[k], b.out_ptr[k + 1]):
                    p = (counts[k, b.out_idx[l]] + transition_pseudocount + b.out_pseudo[l] * use_pseudocount) / norm[k]
                    b.out_logp[l] = log(math.exp(b.out_logp[l]) * edge_inertia + p * (1 - edge_inertia))
            for l in range(b.in_ptr[k], b.in_ptr[k + 1]):
                li = int(b.in_idx[l])
                if norm[li] > 0:
                    p = (counts[li, k] + transition_pseudocount + b.in_pseudo[l] * use_pseudocount) / norm[li]
                    b.in_logp[l] = log(math.exp(b.in_logp[l]) * edge_inertia + p * (1 - edge_inertia))
        visited = numpy.zeros(self.silent_start, dtype=bool)
        for k in range(self.silent_start):                  # once per tied group (4481-4499)
            if visited[k]:
                continue
            visited[k] = True
            for li in b.tied_idx[b.tied_ptr[k]:b.tied_ptr[k + 1]]:
                visited[int(li)] = True
            self.states[k].distribution.from_sample(numpy.asarray(symbols[k]), inertia=distribution_inertia)
        self._push_params()

    # ---- text format (yahmm.pyx:3744-3936) -------------------------------------------------------
    def write(self, stream):
        print("Warning: Writing currently only writes out the model structure,\
			and not any information about tied edges or distributions.")
        b = self._baked
        self.name = self.name.replace(" ", "_")
        stream.write("{} {}\n".format(self.name, len(self.states)))
        for state in sorted(self.states, key=lambda s: s.name):
            state.write(stream)
        for k in range(len(self.states)):
            for l in range(b.out_ptr[k], b.out_ptr[k + 1]):
                li = b.out_idx[l]
                stream.write("{} {} {} {} {} {}\n".format(
                    self.states[k].name.replace(" ", "_"), self.states[li].name.replace(" ", "_"),
                    exp(b.out_logp[l]), b.out_pseudo[l], self.states[k].identity, self.states[li].identity))

    @classmethod
    def read(cls, stream, verbose=False):
        header = stream.readline()
        if header == "":
            raise EOFError("EOF reading HMM header")
        parts = header.strip().split()
        name, num_states = parts[0], int(parts[-1])
        states = {}
        start_state = end_state = None
        for _ in range(num_states):
            state = State.read(stream)
            states[state.identity] = state
            if state.name == "{}-start".format(name):
                start_state = state
            if state.name == "{}-end".format(name):
                end_state = state
        hmm = cls(name=name, start=start_state, end=end_state)
        for state in states.values():
            if state is not start_state and state is not end_state:
                hmm.add_state(state)
        for line in stream:
            if not line.strip():
                continue
            from_name, to_name, prob, pseudo, from_id, to_id = line.strip().split()
            hmm.add_transition(states[from_id], states[to_id], float(prob), float(pseudo))
        hmm.bake(merge=None)
        return hmm

    @classmethod
    def from_matrix(cls, transition_probabilities, distributions, starts, ends, state_names=None, name=None):
        model = cls(name=name)
        states = [State(d, name=n) for n, d in zip(state_names, distributions)]
        n = len(states)
        for state in states:
            model.add_state(state)
        for i, prob in enumerate(starts):
            if prob != 0:
                model.add_transition(model.start, states[i], prob)
        j = 0
        for i in range(n):
            for j, prob in enumerate(transition_probabilities[i]):
                if prob != 0.:
                    model.add_transition(states[i], states[j], prob)
        for i, prob in enumerate(ends):
            if prob != 0:
                # sic: the reference connects states[j] (the LAST column index), not states[i] (yahmm.pyx:3931)
                model.add_transition(states[j], model.end, prob)
        model.bake()
        return model
