"""Approximate (basecall-level) alignment contract (reference nadavca/alignment.py).

``ApproximateSignalAlignment`` and ``signal_alignment_from_base_mapping`` keep the exact layout of
``ApproximateAligner.get_signal_alignment`` (alignment.py:142-186).  Running BWA is host work outside the scope of
this package: any object with ``get_signal_alignment(read, bandwidth)`` can be the aligner.
"""
from collections import namedtuple

import numpy

from .genome import Genome

ApproximateSignalAlignment = namedtuple('ApproximateSignalAlignment',
                                        ['alignment',
                                         'signal_range',
                                         'reference_range',
                                         'read_sequence_range',
                                         'reverse_complement',
                                         'reference_part',
                                         'contig_name'])


def parse_cigar(cigar):
    """'12M3D' -> [(12,'M'),(3,'D')] (alignment.py:42-52)."""
    result, num = [], 0
    for character in cigar:
        if character.isdigit():
            num = num * 10 + int(character)
        else:
            result.append((num, character))
            num = 0
    return result


def base_mapping_from_cigar(cigar, mapped_position, read_sequence, reference, is_reverse_complement):
    """CIGAR walk keeping only matching bases as anchors, flipped to read orientation for the reverse strand
    (alignment.py:109-140)."""
    oriented = Genome.reverse_complement(read_sequence) if is_reverse_complement else numpy.asarray(read_sequence)
    index_in_read, index_in_reference = 0, mapped_position
    mapping = []
    for num, operation in parse_cigar(cigar):
        if operation == 'S' or operation == 'I':
            index_in_read += num
        elif operation == 'D':
            index_in_reference += num
        elif operation == 'M':
            same = numpy.nonzero(numpy.asarray(reference[index_in_reference:index_in_reference + num]) ==
                                 oriented[index_in_read:index_in_read + num])[0]
            mapping.extend((index_in_read + int(i), index_in_reference + int(i)) for i in same)
            index_in_read += num
            index_in_reference += num
        else:
            raise ValueError('Unknown cigar operation: {}'.format(operation))
    if is_reverse_complement:
        mapping = [(len(read_sequence) - 1 - r, len(reference) - 1 - g) for r, g in mapping]
        mapping.reverse()
    return numpy.array(mapping, dtype=int).reshape(-1, 2)


def _mapping_arrays(read):
    """read.sequence_to_signal_mapping (a dict base index -> sample index) as sorted key / value arrays, cached on
    the read (the dict itself is the reference's data contract, read.py:16)."""
    cached = getattr(read, '_mapping_arrays', None)
    mapping = read.sequence_to_signal_mapping
    if cached is None or cached[2] is not mapping or cached[3] != len(mapping):
        keys = numpy.fromiter(mapping.keys(), dtype=numpy.int64, count=len(mapping))
        vals = numpy.fromiter(mapping.values(), dtype=numpy.int64, count=len(mapping))
        order = numpy.argsort(keys, kind='stable')
        cached = (keys[order], vals[order], mapping, len(mapping))
        read._mapping_arrays = cached
    return cached[0], cached[1]


def signal_alignment_from_base_mapping(read, base_mapping, is_reverse_complement, contig_name, reference,
                                       bandwidth):
    """alignment.py:142-186: base anchors -> (signal index, reference index) anchors, ranges and reference part.

    `base_mapping` holds (index in read.sequence, index in the reference) pairs in read orientation -- for the
    reverse strand the reference index counts from the END of the contig, as the reference's CIGAR walk produces."""
    # convert_mapping (alignment.py:58-63): bases the basecaller did not place in the signal are dropped
    base_mapping = numpy.asarray(base_mapping, dtype=int).reshape(-1, 2)
    keys, samples = _mapping_arrays(read)
    pos = numpy.searchsorted(keys, base_mapping[:, 0])
    pos[pos >= len(keys)] = 0
    known = keys[pos] == base_mapping[:, 0] if len(keys) else numpy.zeros(len(base_mapping), dtype=bool)
    signal_mapping = numpy.stack([samples[pos[known]], base_mapping[known, 1]], axis=1).astype(int).reshape(-1, 2)
    if len(signal_mapping) == 0:
        return None
    start_in_reference = signal_mapping[0][1]
    end_in_reference = signal_mapping[-1][1] + 1
    signal_mapping[:, 1] -= start_in_reference
    if is_reverse_complement:
        start_in_reference, end_in_reference = len(reference) - end_in_reference, len(reference) - start_in_reference
    start_in_signal = signal_mapping[0][0]
    end_in_signal = signal_mapping[-1][0] + 1
    extended_start = max(0, start_in_signal - bandwidth)
    extended_end = min(len(read.normalized_signal), end_in_signal + bandwidth)
    signal_mapping[:, 0] -= extended_start
    reference_part = reference[start_in_reference:end_in_reference]
    if is_reverse_complement:
        reference_part = Genome.reverse_complement(reference_part)
    return ApproximateSignalAlignment(alignment=signal_mapping,
                                      signal_range=(extended_start, extended_end),
                                      reference_range=(int(start_in_reference), int(end_in_reference)),
                                      # the UNFILTERED base mapping gives the read range (alignment.py:173-175)
                                      read_sequence_range=(int(base_mapping[0][0]), int(base_mapping[-1][0]) + 1),
                                      reverse_complement=bool(is_reverse_complement),
                                      reference_part=reference_part,
                                      contig_name=contig_name)


_CIGAR_CODES = {'M': 0, 'I': 1, 'D': 2, 'S': 3}


def batch_signal_alignments(reads, hits, reference, bandwidth, device=0, contig_name='', d_reference=None):
    """The anchor construction of ``ApproximateAligner.get_signal_alignment`` (alignment.py:109-186) for a whole
    batch on the GPU (csrc/anchors.cu): `hits[i]` is the BWA hit of `reads[i]` as ``(cigar, mapped_position,
    is_reverse_complement)`` -- the three fields the reference takes from the SAM record (alignment.py:100-107) -- or
    None for an unmapped read.  Returns one ``ApproximateSignalAlignment`` (or None) per read, equal field by field
    to ``signal_alignment_from_base_mapping(read, base_mapping_from_cigar(...), ...)``.  `d_reference`: the contig's
    numeric codes already resident on the device (a torch int8 tensor) to skip its upload."""
    import ctypes
    import re
    import torch
    from . import _cabi
    lib = _cabi.require_device()
    idx = [i for i, h in enumerate(hits) if h is not None]
    results = [None] * len(reads)
    if not idx:
        return results
    dev = torch.device('cuda', int(device))
    if d_reference is None:
        d_reference = torch.as_tensor(reference_codes(reference), device=dev)
    n = len(idx)
    cig_len, cig_op, cig_off = [], [], numpy.zeros(n + 1, dtype=numpy.int64)
    read_off = numpy.zeros(n + 1, dtype=numpy.int64)
    seqs, maps = [], []
    pos = numpy.zeros(n, dtype=numpy.int64)
    rev = numpy.zeros(n, dtype=numpy.int32)
    n_signal = numpy.zeros(n, dtype=numpy.int32)
    for j, i in enumerate(idx):
        cigar, mapped_position, is_rc = hits[i][:3]
        ops = re.findall(r'(\d+)(.)', cigar)
        for num, op in ops:
            if op not in _CIGAR_CODES:
                raise ValueError('Unknown cigar operation: {}'.format(op))
            cig_len.append(int(num))
            cig_op.append(_CIGAR_CODES[op])
        cig_off[j + 1] = len(cig_len)
        read = reads[i]
        seq = reference_codes(read.sequence)
        keys, samples = _mapping_arrays(read)
        dense = numpy.full(len(seq), -1, dtype=numpy.int32)
        ok = (keys >= 0) & (keys < len(seq))
        dense[keys[ok]] = samples[ok]
        seqs.append(seq)
        maps.append(dense)
        read_off[j + 1] = read_off[j] + len(seq)
        pos[j], rev[j], n_signal[j] = mapped_position, int(bool(is_rc)), len(read.normalized_signal)
    cig_len = numpy.ascontiguousarray(cig_len, dtype=numpy.int32)
    cig_op = numpy.ascontiguousarray(cig_op, dtype=numpy.int8)
    seq_all = numpy.ascontiguousarray(numpy.concatenate(seqs), dtype=numpy.int8)
    map_all = numpy.ascontiguousarray(numpy.concatenate(maps), dtype=numpy.int32)
    anchors = numpy.zeros(2 * int(read_off[-1]), dtype=numpy.int32)
    meta = numpy.zeros(7 * n, dtype=numpy.int64)
    h = _cabi.NvbHits(n, _cabi.ptr(cig_len, ctypes.c_int32), _cabi.ptr(cig_op, ctypes.c_int8),
                      _cabi.ptr(cig_off, ctypes.c_int64), _cabi.ptr(pos, ctypes.c_int64), _cabi.ptr(rev, ctypes.c_int32),
                      _cabi.ptr(seq_all, ctypes.c_int8), _cabi.ptr(map_all, ctypes.c_int32),
                      _cabi.ptr(read_off, ctypes.c_int64), _cabi.ptr(n_signal, ctypes.c_int32),
                      ctypes.c_void_p(d_reference.data_ptr()), len(reference), int(bandwidth))
    stream = torch.cuda.current_stream(dev)
    _cabi.check(lib.nvb_signal_anchors_batch(int(device), ctypes.byref(h), _cabi.ptr(anchors, ctypes.c_int32),
                                             _cabi.ptr(meta, ctypes.c_int64), ctypes.c_void_p(stream.cuda_stream)),
                'nvb_signal_anchors_batch')
    meta = meta.reshape(n, 7)
    for j, i in enumerate(idx):
        count, ref_start, ref_end, sig_start, sig_end, read_start, read_end = (int(x) for x in meta[j])
        if count == 0:
            continue
        is_rc = bool(rev[j])
        part = reference[ref_start:ref_end]
        if is_rc:
            part = Genome.reverse_complement(part)
        pairs = anchors[2 * read_off[j]:2 * read_off[j] + 2 * count].reshape(count, 2).astype(int)
        results[i] = ApproximateSignalAlignment(alignment=pairs, signal_range=(sig_start, sig_end),
                                                reference_range=(ref_start, ref_end),
                                                read_sequence_range=(read_start, read_end), reverse_complement=is_rc,
                                                reference_part=part,
                                                contig_name=hits[i][3] if len(hits[i]) > 3 else contig_name)
    return results


_CODE_LUT = numpy.full(256, 4, dtype=numpy.int8)
for _code, _base in enumerate('ACGT'):
    _CODE_LUT[ord(_base)] = _code


def reference_codes(sequence):
    """Bases -> int8 codes 0..3 (4 for anything else); accepts a str, an array of 1-char strings or of codes."""
    if isinstance(sequence, str):
        return _CODE_LUT[numpy.frombuffer(sequence.encode('ascii'), dtype=numpy.uint8)]
    arr = numpy.asarray(sequence)
    if arr.dtype.kind in 'iu':
        return arr.astype(numpy.int8)
    if arr.size == 0:
        return numpy.zeros(0, dtype=numpy.int8)
    from .genome import _as_bytes
    return _CODE_LUT[_as_bytes(arr)]


class ApproximateAligner:
    """Placeholder with the reference's constructor (alignment.py:19-40).  Mapping reads with BWA is host work that
    stays outside this package; plug in any aligner object exposing ``get_signal_alignment(read, bandwidth)``
    (``signal_alignment_from_base_mapping`` builds its return value from a CIGAR-derived base mapping)."""

    def __init__(self, bwa_executable, reference, reference_filename, references_dict=None):
        self.bwa_executable = bwa_executable
        self.reference = reference
        self.references_dict = references_dict
        self.reference_filename = reference_filename

    def get_signal_alignment(self, read, bandwidth):
        raise NotImplementedError('BWA mapping is out of scope for nadavca_b200: pass an `aligner` object that '
                                  'implements get_signal_alignment(read, bandwidth)')
