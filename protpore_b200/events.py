"""Synthetic nanopore events (host side): length-conditioned ancestral sampling from a baked model.

`Model.sample(length=L)` of the reference is unreliable on finite models (yahmm.pyx:2742-2771 reuses a
stale loop index), and free-running samples of the pro_001 model average only 15 segments, so the
benchmark events (50-500, 10k, 20-20,000 segments) come from this sampler: walk the baked
out-edge CSR from `start`; while fewer than L symbols are out, edges into `end` -- and into states that
can no longer reach an emitting state -- are masked and the remaining probabilities renormalised; stop
after the L-th emission.  Values are rounded to float32, which is what the GPU path stores.

This is the small-scale host twin of the CUDA sampler behind `Engine.sample_events` (k_sample); the two
use different random streams and are not expected to produce identical events.
"""
import numpy as np


def _alive(baked):
    m, ne = len(baked.states), baked.silent_start
    alive = np.zeros(m, dtype=bool)
    alive[:ne] = True
    changed = True
    while changed:
        changed = False
        for k in range(ne, m):
            if alive[k] or k == baked.end_index:
                continue
            dst = baked.out_idx[baked.out_ptr[k]:baked.out_ptr[k + 1]]
            if any(d != baked.end_index and alive[d] for d in dst):
                alive[k] = True
                changed = True
    return alive


def sample_events(model, lengths, seed=0):
    """List of float32 arrays, event e having exactly lengths[e] observations."""
    b = model._baked
    rng = np.random.RandomState(seed)
    alive = _alive(b)
    probs = np.exp(b.out_logp)
    params = [s.distribution.device_params() for s in b.states[:b.silent_start]]
    out = []
    for length in lengths:
        while True:
            xs, s, dead = [], b.start_index, False
            while len(xs) < length:
                if s < b.silent_start:
                    kind, p0, p1 = params[s]
                    xs.append(rng.normal(p0, p1) if (kind == 0 and p1 > 0) else (p0 if kind == 0 else rng.uniform(p0, p1)))
                    if len(xs) == length:
                        break
                lo, hi = b.out_ptr[s], b.out_ptr[s + 1]
                dst = b.out_idx[lo:hi]
                w = np.where((dst != b.end_index) & alive[dst], probs[lo:hi], 0.0)
                if w.sum() <= 0:
                    dead = True
                    break
                s = int(dst[min(int(np.searchsorted(np.cumsum(w), rng.uniform(0, w.sum()), side='right')), hi - lo - 1)])
            if not dead:
                break
        out.append(np.asarray(xs, dtype=np.float32))
    return out


def uniform_lengths(n_events, lo, hi, seed=0):
    """UniformInt[lo, hi] event lengths (BASELINE config 2: 50-500)."""
    return np.random.RandomState(seed).randint(lo, hi + 1, size=n_events).astype(np.int64)


def log_uniform_lengths(n_events, lo, hi, seed=0):
    """Log-uniform event lengths (BASELINE config 5: 20-20,000)."""
    u = np.random.RandomState(seed).uniform(np.log(lo), np.log(hi), size=n_events)
    return np.clip(np.round(np.exp(u)), lo, hi).astype(np.int64)
