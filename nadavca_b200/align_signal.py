"""align_signal() -- reference nadavca/align_signal.py:13-81, batched on the GPU."""
import sys

import numpy as np
from scipy.stats import linregress

from . import defaults
from .alignment import ApproximateAligner
from .estimator import ProbabilityEstimator
from .genome import Genome
from .kmer_model import KmerModel
from .read import Read


def load_model_and_estimator(reference_filename, config=defaults.CONFIG_FILE, kmer_model=None,
                             bwa_executable=defaults.BWA_EXECUTABLE, aligner=None, reference=None):
    """align_signal.py:13-41 -> (kmer_model, ProbabilityEstimator) or None when a file is missing."""
    if kmer_model is None:
        kmer_model = defaults.KMER_MODEL_FILE
    try:
        config = defaults.load_config(config)
    except FileNotFoundError:
        sys.stderr.write('failed to load config: {} not found\n'.format(config))
        return None
    if isinstance(kmer_model, str):
        kmer_model = KmerModel.load_from_hdf5(kmer_model)
    if aligner is None:
        try:
            references = Genome.load_from_fasta(reference_filename)
        except FileNotFoundError:
            sys.stderr.write("failed to process: reference {} doesn't exist\n".format(reference_filename))
            return None
        references_dict = {r.description[1:]: r.bases for r in references}
        aligner = ApproximateAligner(bwa_executable, reference, reference_filename, references_dict)
    return kmer_model, ProbabilityEstimator(kmer_model, aligner, config)


def _linear_renormalization(kmer_model, read, apx_alignment, alignment, signal_means=None, expected=None):
    """One even renorm round (align_signal.py:59-76): regress per-event means on the expected levels and rescale
    the whole normalised signal.  `signal_means` are the device-computed event means (bit-identical to the
    reference's per-event numpy.mean loop) and `expected` the expected levels with EMPTY contexts
    (align_signal.py:63); without them both are computed here, one read at a time."""
    if expected is None:
        bases_num = Genome.to_numerical(apx_alignment.reference_part)
        expected = np.array(kmer_model.get_expected_signal(bases_num, [], []))
    if signal_means is None:
        signal_cut = read.normalized_signal[alignment[0][1]:alignment[-1][2]]
        al_start = alignment[0][1]
        signal_means = [np.mean(signal_cut[s - al_start:e - al_start]) for _, s, e in alignment]
    slope, intercept = linear_fit(expected, signal_means)
    read.normalized_signal = (read.normalized_signal - intercept) / slope


def linear_fit(x, y):
    """slope and intercept of scipy.stats.linregress(x, y) (align_signal.py:73) without its p-value / standard-error
    machinery, which costs ~0.5 ms per read.  Same numpy calls in the same order as scipy 1.18's implementation
    (mean, demean, vecdot / n), so the two numbers are bit-identical to scipy's --
    tests/test_host_logic.py::test_linear_fit_equals_scipy_linregress fails loudly if an installed scipy computes
    them differently; anything unusual (masked arrays, fewer than 2 points, constant x) goes to scipy itself."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    n = x.size
    if x.ndim != 1 or y.shape != x.shape or n < 2 or not hasattr(np, 'vecdot') or \
            not (np.isfinite(x).all() and np.isfinite(y).all()):
        res = linregress(x, y)
        return res.slope, res.intercept
    xmean = np.mean(x, axis=-1, keepdims=True)
    ymean = np.mean(y, axis=-1, keepdims=True)
    x_ = x - xmean
    y_ = y - ymean
    ssxm = np.vecdot(x_, x_, axis=-1) / n
    ssxym = np.vecdot(x_, y_, axis=-1) / n
    if ssxm == 0.0:
        res = linregress(x, y)
        return res.slope, res.intercept
    slope = ssxym / ssxm
    intercept = ymean[0] - slope * xmean[0]
    return slope, intercept


def align_signal(reference_filename,
                 reads,
                 config=defaults.CONFIG_FILE,
                 kmer_model=defaults.KMER_MODEL_FILE,
                 bwa_executable=defaults.BWA_EXECUTABLE,
                 group_name=defaults.GROUP_NAME,
                 renorm_rounds=defaults.RENORM_ROUNDS,
                 aligner=None,
                 reference=None):
    """Generator of ``(read, (approximate_alignment, alignment))`` like the reference (align_signal.py:43-81);
    `alignment` is an int (n,3) array [reference position, event_start, event_end].  Accepts ``Read`` instances
    (fast5 file names need h5py and are out of scope).  Reads without an alignment yield ``(read, None)`` as the
    reference's README promises.  All reads go through each stage as one GPU batch: refine -> linear
    renormalisation (round 0) -> refine (round 1) -> linear renormalisation (round 2)."""
    if aligner is None:
        raise ValueError('align_signal needs an `aligner` (any object with get_signal_alignment(read, bandwidth)): '
                         'mapping reads with BWA is host work outside nadavca_b200')
    loaded = load_model_and_estimator(reference_filename, config, kmer_model, bwa_executable, aligner, reference)
    if loaded is None:
        return
    kmer_model, estimator = loaded
    reads = list(reads)
    for i, read in enumerate(reads):
        if isinstance(read, str):
            reads[i] = Read.load_from_fast5(read, group_name)
    Read.normalize_each(reads, kmer_model.device)  # per-read median / MAD (align_signal.py:54), one batch on the GPU
    results = estimator.get_refined_alignments(reads, with_event_means=True)
    prepared = estimator.last_prepared  # the approximate alignments do not change with the normalisation
    for r in range(renorm_rounds):
        alive = [i for i, res in enumerate(results) if res is not None]
        if r % 2 == 0:
            refs = [Genome.to_numerical(results[i][0].reference_part) for i in alive]
            expected = kmer_model.get_expected_signal_batch(refs, [[]] * len(alive), [[]] * len(alive))
            for i, exp in zip(alive, expected):
                _linear_renormalization(kmer_model, reads[i], *results[i], expected=exp)
        else:
            again = estimator.get_refined_alignments([reads[i] for i in alive], with_event_means=True,
                                                     prepared=[prepared[i] for i in alive])
            for i, res in zip(alive, again):
                results[i] = res
    for read, res in zip(reads, results):
        yield read, (None if res is None else res[:2])
