"""nadavca_b200 -- B200-native implementation of nadavca's data-parallel hot path.

Public surface mirrors the reference package (nadavca/__init__.py:1-2 plus the classes its README documents).
"""
from .estimate_snps import estimate_snps  # noqa: F401
from .align_signal import align_signal  # noqa: F401
from .detect_meth import detect_meth  # noqa: F401
from .read import Read  # noqa: F401
from .kmer_model import KmerModel  # noqa: F401
from . import estimator, dtw, alignment, genome, synthetic  # noqa: F401
