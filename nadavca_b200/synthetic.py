"""Seeded synthetic genomes / reads simulated from a k-mer model, and the aligner that stands in for BWA.

Generator specification: SURVEY.md 8(d).  The ``SyntheticAligner`` plugs into the ``aligner`` seam of
``ProbabilityEstimator`` (reference nadavca/estimator.py:34-36,60,159) and builds its result with the same layout
as ``ApproximateAligner.get_signal_alignment`` (alignment.py:142-186).  Host-side numpy only.
"""
import numpy as np

from .alignment import signal_alignment_from_base_mapping
from .alphabet import alphabet
from .genome import Genome
from .read import Read

_ALPHA = np.array(alphabet)


def make_genome(length, seed=0):
    """Uniform i.i.d. ACGT as an array of 1-char strings."""
    rng = np.random.default_rng(seed)
    return _ALPHA[rng.integers(0, 4, size=length)]


def kmer_ids(numeric, k, central):
    """k-mer id at every position of `numeric` with base-0 ('A') padding outside (sequence.cpp:23-28)."""
    n = len(numeric)
    padded = np.zeros(n + k, dtype=np.int64)
    padded[central:central + n] = numeric
    ids = np.zeros(n, dtype=np.int64)
    for j in range(k):
        ids = ids * 4 + padded[j:j + n]
    return ids


def make_read(genome, kmer_model, index, n_bases=None, bandwidth=150, flank=8, substitution_rate=0.0,
              mean_event=10, jitter=20, seed_base=1000, strand=None, start=None):
    """One synthetic read (seed = seed_base + index) drawn from `genome`; returns a ``Read`` carrying its truth."""
    rng = np.random.default_rng(seed_base + index)
    G = len(genome)
    if n_bases is None:
        n_bases = int(round(rng.normal(2000, 200)))
    n = int(max(8, min(n_bases, G)))
    if start is None:
        start = int(rng.integers(0, G - n + 1))
    reverse = bool(rng.integers(0, 2)) if strand is None else (strand == '-')
    segment = genome[start:start + n]
    oriented = Genome.reverse_complement(segment) if reverse else segment
    flank_l = _ALPHA[rng.integers(0, 4, size=flank)]
    flank_r = _ALPHA[rng.integers(0, 4, size=flank)]
    molecule = np.concatenate([flank_l, oriented, flank_r])
    numeric = Genome.to_numerical(molecule)
    k, cp = kmer_model.get_k(), kmer_model.get_central_position()
    levels = kmer_model.mean[kmer_ids(numeric, k, cp)]
    sd = kmer_model.sigma[kmer_ids(numeric, k, cp)]
    lengths = np.maximum(2, rng.poisson(mean_event, size=len(molecule)))
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]) + bandwidth
    body = np.repeat(levels, lengths) + rng.normal(0, 1, size=int(lengths.sum())) * np.repeat(sd, lengths)
    body = np.clip(body, -5, 5)
    level = np.concatenate([rng.normal(0, 1, size=bandwidth), body, rng.normal(0, 1, size=bandwidth)])
    raw = 15.0 * level + 90.0

    sequence = molecule.copy()
    if substitution_rate > 0:
        flips = np.nonzero(rng.random(len(sequence)) < substitution_rate)[0]
        for pos in flips:
            sequence[pos] = _ALPHA[(Genome.to_numerical(sequence[pos:pos + 1])[0] + rng.integers(1, 4)) % 4]
    noisy = starts + rng.integers(-jitter, jitter + 1, size=len(starts))
    noisy = np.clip(np.maximum.accumulate(noisy), 0, len(raw) - 1)
    read = Read.from_arrays(raw, sequence, {int(b): int(s) for b, s in enumerate(noisy)},
                            name='synthetic_{}'.format(index))
    read.truth = {'start': start, 'n': n, 'reverse': reverse, 'flank': flank, 'event_starts': starts,
                  'event_lengths': lengths, 'molecule': molecule}
    return read


def make_reads(genome, kmer_model, count, **kwargs):
    return [make_read(genome, kmer_model, i, **kwargs) for i in range(count)]


class SyntheticAligner:
    """Stands in for BWA: the base-level mapping is known from the simulation (``read.truth``)."""

    def __init__(self, reference, contig_name='synthetic'):
        self.reference = reference
        self.contig_name = contig_name

    def get_signal_alignment(self, read, bandwidth):
        truth = getattr(read, 'truth', None)
        if truth is None:
            return None
        G = len(self.reference)
        start, n, flank = truth['start'], truth['n'], truth['flank']
        segment = self.reference[start:start + n]
        oriented = Genome.reverse_complement(segment) if truth['reverse'] else segment
        read_idx = np.arange(n) + flank
        same = np.nonzero(read.sequence[read_idx] == oriented)[0]  # only matching bases anchor (alignment.py:122-124)
        if truth['reverse']:
            # reference indices count from the end of the contig in read orientation (alignment.py:134-138)
            ref_idx = (G - (start + n)) + same
        else:
            ref_idx = start + same
        base_mapping = np.stack([read_idx[same], ref_idx], axis=1).astype(int)
        if len(base_mapping) == 0:
            return None
        return signal_alignment_from_base_mapping(read, base_mapping, truth['reverse'], self.contig_name,
                                                  self.reference, bandwidth)
