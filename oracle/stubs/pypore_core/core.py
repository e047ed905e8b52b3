"""Stand-in for PyPore/core.py while the compiled reference cparsers module is imported (TEST INFRASTRUCTURE).

cparsers.pyx:17 does `from core import Segment` and only constructs Segment(current=..., start=..., duration=...,
end=...) (cparsers.pyx:115-116); PyPore/core.py itself is Python-2 source and cannot be imported here."""
import numpy as np


class Segment(object):
    def __init__(self, current=None, **kwargs):
        self.current = current
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def mean(self):                      # PyPore/core.py:209-211
        return np.mean(self.current)

    @property
    def std(self):                       # PyPore/core.py:212-214
        return np.std(self.current)
