"""`State` of the drop-in yahmm API (yahmm.pyx:1920-2019)."""
from .distributions import *  # noqa: F401,F403  (State.read evaluates distribution expressions)


class State(object):
    """A node of the HMM graph: an emission distribution (None = silent), a name, an identity, a weight."""

    def __init__(self, distribution, name=None, weight=None, identity=None):
        self.distribution = distribution
        self.name = name or str(id(str))          # sic: the reference uses id(str) (yahmm.pyx:1946)
        self.identity = str(identity) if identity is not None else str(id(self))
        self.weight = weight or 1.

    def __str__(self):
        return "State( {}, name={}, weight={}, identity={} )".format(
            str(self.distribution), self.name, self.weight, self.identity)

    def is_silent(self):
        return self.distribution is None

    def tied_copy(self):
        return State(distribution=self.distribution, name=self.name)

    def copy(self):
        return State(self.distribution, name=self.name, weight=self.weight, identity=self.identity)

    def write(self, stream):
        stream.write("{} {} {} {}\n".format(
            self.identity, self.name.replace(" ", "_"), self.weight, str(self.distribution)))

    @classmethod
    def read(cls, stream):
        line = stream.readline()
        if line == "":
            raise EOFError("End of file while reading state.")
        parts = line.strip().split()
        identity, name, weight, distribution = parts[0], parts[1], parts[2], ' '.join(parts[3:])
        return eval("State( {}, name='{}', weight={}, identity='{}' )".format(
            distribution, name, weight, identity))
