"""Shared helpers of the test-suite: golden fixtures, model builders, oracle construction."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

_spec = importlib.util.spec_from_file_location("make_golden_builders", os.path.join(GOLDEN, "make_golden.py"))

MODEL_NAMES = ["pro_001_py2", "pro_001_py3", "synth_12", "toy", "toy_uniform", "toy_infinite"]


def builders():
    """name -> callable(yahmm_module) building the fixture's model (same code the golden script used)."""
    from protpore_b200 import peptide_hmm
    import ast
    src = open(os.path.join(GOLDEN, "make_golden.py")).read()
    tree = ast.parse(src)
    ns = {"peptide_hmm": peptide_hmm}
    for node in tree.body:      # pull out only the toy-model builder functions
        if isinstance(node, ast.FunctionDef) and node.name.startswith("toy"):
            exec(compile(ast.Module([node], []), "make_golden.py", "exec"), ns)
    return {
        "pro_001_py2": lambda y: peptide_hmm.pro_001_model(yahmm_module=y, py2_division=True),
        "pro_001_py3": lambda y: peptide_hmm.pro_001_model(yahmm_module=y, py2_division=False),
        "synth_12": lambda y: peptide_hmm.model_maker(peptide_hmm.synthetic_profile(12, seed=3), "synth12", y, True),
        "toy": ns["toy"], "toy_uniform": ns["toy_uniform"], "toy_infinite": ns["toy_infinite"],
    }


def build_model(name):
    import protpore_b200.yahmm as y
    return builders()[name](y)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_events(g):
    off = g["offsets"]
    return [g["obs"][off[i]:off[i + 1]] for i in range(len(off) - 1)]


def golden_paths(g):
    off = g["path_offsets"]
    return [g["paths"][off[i]:off[i + 1]] if np.isfinite(g["vscore"][i]) else None for i in range(len(off) - 1)]


def oracle_for(model):
    from oracle import cpu_oracle
    return cpu_oracle.OracleModel.from_model(model)


def close(a, b, abs_tol=1e-3, rel_tol=1e-6):
    """The north-star tolerance: |d| <= 1e-3 nats absolute OR 1e-6 relative; infinities must agree."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    inf = ~np.isfinite(a) | ~np.isfinite(b)
    ok = np.zeros(a.shape, dtype=bool)
    ok[inf] = (a[inf] == b[inf]) | (np.isnan(a[inf]) & np.isnan(b[inf]))
    d = np.abs(a[~inf] - b[~inf])
    ok[~inf] = (d <= abs_tol) | (d <= rel_tol * np.abs(b[~inf]))
    return ok


def normalise_identities(text):
    """Model.write output with every state identity (id()-based in the reference, yahmm.pyx:1948) replaced by the ordinal
    of its first appearance, so that two runs -- or two implementations -- can be compared byte for byte."""
    lines = text.splitlines()
    n_states = int(lines[0].split()[-1])
    ids = {}
    out = [lines[0]]
    for ln in lines[1:1 + n_states]:
        parts = ln.split(" ")
        ids.setdefault(parts[0], "id%d" % len(ids))
        out.append(" ".join([ids[parts[0]]] + parts[1:]))
    for ln in lines[1 + n_states:]:
        parts = ln.split(" ")
        out.append(" ".join(parts[:-2] + [ids[parts[-2]], ids[parts[-1]]]))
    return "\n".join(out) + "\n"


def _dist_from_spec(y, spec):
    t = spec["type"]
    if t in ("GaussianKernelDensity", "UniformKernelDensity", "TriangleKernelDensity"):
        return getattr(y, t)(spec["points"], spec["bandwidth"], spec["weights"])
    if t == "MixtureDistribution":
        return y.MixtureDistribution([_dist_from_spec(y, p) for p in spec["parts"]], spec["weights"])
    return getattr(y, t)(*spec["parameters"])


def replay_model(log):
    """Rebuild, on this repository's yahmm, the graph a reference builder constructed (tests/golden/make_golden_pypore.py
    recorded it node by node and edge by edge in insertion order), and bake it."""
    import protpore_b200.yahmm as y
    dists = [_dist_from_spec(y, d) for d in log["dists"]]
    nodes = log["nodes"]
    states = [y.State(None if n["dist"] is None else dists[n["dist"]], name=n["name"], weight=n["weight"]) for n in nodes]
    model = y.Model(name=log["name"], start=states[log["start"]], end=states[log["end"]])
    from protpore_b200.yahmm import _graph
    g = _graph.OrderedDiGraph()
    for s in states:
        g.add_node(s)
    for a, b, w, pc, grp in log["edges"]:
        g.add_edge(states[a], states[b], weight=w, pseudocount=pc, group=grp)
    model.graph = g
    model.bake()
    return model
