"""Import stub: the reference's yahmm.pyx imports matplotlib.pyplot (yahmm.pyx:28) only for Model.draw.

Test infrastructure only (see oracle/README.md)."""
rcParams = {}
