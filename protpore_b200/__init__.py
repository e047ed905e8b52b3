"""protpore_b200 -- B200-native scoring / decoding of segmented nanopore events against profile HMMs.

Public surface (mirrors the one hot path of mitenjain/protpore):
  protpore_b200.yahmm        drop-in `yahmm` (Model / State / distributions), sweeps on the GPU
  protpore_b200.peptide_hmm  `HMM_Constructor`, `model_maker`, `prediction`, `kmer_current_map`
  protpore_b200.pypore       `Segment`, `Event.apply_hmm`, `apply_hmm_batch`
  protpore_b200.parsers      `SpeedyStatSplit` / `FastStatSplit` (raw current -> segments), `parse_batch`
  protpore_b200.parallel     one-process-per-GPU sharding + NCCL allreduce for Baum-Welch
"""
import sys

__version__ = "0.1.0"


def install_as_yahmm():
    """Register `protpore_b200.yahmm` under the name `yahmm`, so `import yahmm` / `from yahmm import *`
    in PyPore and peptide_hmm_v0.0.1.py resolve to the GPU-backed implementation."""
    from . import yahmm
    sys.modules["yahmm"] = yahmm
    return yahmm
