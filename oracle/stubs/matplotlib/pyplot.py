"""Import stub for `from matplotlib import pyplot` (yahmm.pyx:28). Only Model.draw (2247-2259) uses it."""


def show(*args, **kwargs):
    pass
