"""detect_meth() -- reference nadavca/detect_meth.py:21-110 on top of the batched GPU alignment.

The reference aligns one read at a time, renormalises, re-aligns and then scores every occurrence of `pattern` in the
aligned reference part: for the 11 positions around the occurrence, the mean level of the base's event is compared with
the k-mer model's expected level (two-sided normal p-value with a fixed sd, reported as -log p).  Here all reads go
through ``align_signal`` (batched refine / renormalise rounds on the device); the scoring itself is a few thousand
numpy / scipy operations per read and stays on the host with the reference's own calls.
"""
import csv
import sys

import numpy as np
from scipy import stats

from . import defaults
from .align_signal import align_signal
from .genome import Genome

SMALLEST_PVAL = 1e-50
LEVEL_SD = 0.35287208  # detect_meth.py:22
WINDOW = 5             # positions on either side of the pattern start (detect_meth.py:46)


def cdf_scoring(raw, exp):
    """-log of the two-sided p-value of the event mean under N(exp, LEVEL_SD) (detect_meth.py:21-24)."""
    z_score = np.abs(raw.mean() - exp) / LEVEL_SD
    p_value = stats.norm.cdf(-z_score) * 2.0
    return -np.log(max(SMALLEST_PVAL, p_value))


def calculate_meth_scores(signal_cut, alignment, apx_alignment, pattern, kmer_model):
    """detect_meth.py:26-55 -> list of (position, 11-base context, 11 scores) for every occurrence of `pattern` whose
    whole +-5 window lies inside the alignment and has no empty event.  `signal_cut` starts at alignment[0][1]."""
    bases = apx_alignment.reference_part
    expected = np.array(kmer_model.get_expected_signal(Genome.to_numerical(bases), [], []))  # empty contexts (Q6)
    sequence = ''.join(bases)
    origin = alignment[0][1]
    features = []
    pos = sequence.find(pattern)
    while pos != -1:
        scores = []
        for i in range(pos - WINDOW, pos + WINDOW + 1):
            if i < 0 or i >= len(alignment):
                continue
            event = signal_cut[alignment[i][1] - origin:alignment[i][2] - origin]
            if len(event) == 0:
                break
            scores.append(cdf_scoring(event, expected[i]))
        if len(scores) == 2 * WINDOW + 1:
            # the reference slices seq[pos-5:pos+6]; a full window implies pos >= 5, so no negative index arises
            features.append((pos, sequence[pos - WINDOW:pos + WINDOW + 1], scores))
        pos = sequence.find(pattern, pos + 1)
    return features


def maxs3(values):
    """Largest sum of three consecutive scores (detect_meth.py:58-60)."""
    return max(a + b + c for a, b, c in zip(values, values[1:], values[2:]))


def detect_meth(reference_filename,
                reads,
                pattern,
                output,
                config=defaults.CONFIG_FILE,
                kmer_model=defaults.KMER_MODEL_FILE,
                bwa_executable=defaults.BWA_EXECUTABLE,
                group_name=defaults.GROUP_NAME,
                renorm_rounds=defaults.RENORM_ROUNDS,
                aligner=None,
                reference=None,
                names=None):
    """Same arguments as the reference (detect_meth.py:63-110) plus the `aligner` / `reference` seams of
    ``align_signal`` and optional row labels `names` (the reference writes the fast5 file name).  Writes the CSV
    (header + one row per scored pattern occurrence) to `output`, or to stdout when it is None, and returns the rows.
    Reads without an alignment are skipped (the reference crashes on them)."""
    from .kmer_model import KmerModel
    if isinstance(kmer_model, str):
        kmer_model = KmerModel.load_from_hdf5(kmer_model)
    reads = list(reads)
    rows = []
    aligned = align_signal(reference_filename, reads, config, kmer_model, bwa_executable, group_name, renorm_rounds,
                           aligner=aligner, reference=reference)
    for index, (read, result) in enumerate(aligned):
        if result is None:
            continue
        apx_alignment, alignment = result
        signal_cut = read.normalized_signal[alignment[0][1]:alignment[-1][2]]
        label = names[index] if names is not None else (reads[index] if isinstance(reads[index], str) else index)
        for pos, context, scores in calculate_meth_scores(signal_cut, alignment, apx_alignment, pattern, kmer_model):
            rows.append((label, pos, context, ','.join(map(str, scores)), maxs3(scores)))
    out = open(output, 'w') if output is not None else sys.stdout
    try:
        writer = csv.writer(out)
        writer.writerow(('Filename', 'Position', 'Sequence context', 'Position scores', 'Aggregated score'))
        writer.writerows(rows)
    finally:
        if output is not None:
            out.close()
    return rows
