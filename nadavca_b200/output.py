"""On-disk outputs of the reference's CLI commands (host I/O):

  * the per-read ``.npz`` of ``nadavca align`` (align_signal.py:83-132): ``arr_0`` = the raw signal between the first
    and the last event start, ``arr_1`` = one label per sample ('N', or the reference base at the sample where its
    event starts), ``arr_2`` = [reference start, strand, contig name, read sequence];
  * the SNP table of ``nadavca snp`` (estimate_snps.py:97-127, estimator.py:21-31): ``index base coverage A C G T``
    with probabilities printed as ``{:18.16f}``; one file per read in independent mode, one table otherwise.
"""
import os
import sys

import numpy as np

from .estimator import Chunk


class AlignException(Exception):
    """Raised for an alignment whose event starts are not strictly increasing (align_signal.py:127 names this
    exception without defining it)."""


def sample_labels(apx_alignment, alignment):
    """align_signal.py:116-130: per-sample base labels over raw_signal[alignment[0][1]:alignment[-1][1]]."""
    alignment = np.asarray(alignment)
    first = int(alignment[0][1])
    starts = alignment[:-1, 1]
    if len(starts) and (starts[0] <= -47 or np.any(np.diff(starts) <= 0)):
        raise AlignException('bad alignment')
    labels = np.full(int(alignment[-1][1]) - first, 'N')
    if len(labels):
        labels[starts - first] = np.asarray(apx_alignment.reference_part)[:len(starts)]
    return labels


def write_alignment_npz(filename, read, apx_alignment, alignment):
    """Write ``filename + '.npz'`` like write_binary_output (align_signal.py:83-84); returns the path or None for an
    empty cut (the reference prints "empty raw cut" and skips the read)."""
    alignment = np.asarray(alignment)
    raw_cut = np.asarray(read.raw_signal)[int(alignment[0][1]):int(alignment[-1][1])]
    if len(raw_cut) == 0:
        return None
    meta = np.array([str(apx_alignment.reference_range[0]), '-' if apx_alignment.reverse_complement else '+',
                     str(apx_alignment.contig_name), ''.join(read.sequence)])
    np.savez(filename + '.npz', raw_cut, sample_labels(apx_alignment, alignment), meta)
    return filename + '.npz'


def write_alignments(results, output_dir, names):
    """The writer loop of align_signal_command (align_signal.py:100-132) over ``align_signal()`` results."""
    os.makedirs(output_dir, exist_ok=True)
    written = []
    for (read, res), name in zip(results, names):
        if res is None:
            continue
        base = os.path.splitext(os.path.basename(name))[0]
        path = write_alignment_npz(os.path.join(output_dir, base), read, res[0], res[1])
        if path is None:
            print('empty raw cut', name)
        else:
            written.append(path)
    return written


def write_snp_tables(chunks, reference, output=None, independent=False, names=None):
    """estimate_snps_command's writer (estimate_snps.py:97-127).  Independent mode: one ``<name>.txt`` per read in
    directory `output` (stdout when `output` is None); consensus mode: one table in file `output`."""
    if independent:
        if output:
            os.makedirs(output, exist_ok=True)
        names = names if names is not None else ['read_{}'.format(i) for i in range(len(chunks))]
        for chunk, name in zip(chunks, names):
            if chunk is None:
                continue
            fh = open(os.path.join(output, os.path.splitext(os.path.basename(name))[0] + '.txt'), 'w') \
                if output else sys.stdout
            Chunk.print_head(fh)
            chunk.print(fh, reference)
            if output:
                fh.close()
        return
    fh = open(output, 'w') if output else sys.stdout
    Chunk.print_head(fh)
    for chunk in chunks:
        chunk.print(fh, reference)
    if output:
        fh.close()
