"""Read container and host-side signal normalisation (reference nadavca/read.py).

fast5 parsing needs h5py / ont formats and is out of scope; reads are built with
``Read.from_arrays`` or by ``nadavca_b200.synthetic``.  Normalisation and the spline tweak stay on the host and use
the same numpy / scipy calls as the reference, so the kernels see identical inputs.
"""
import numpy
from scipy import interpolate


class Read:
    def __init__(self):
        self.raw_signal = None
        self.normalized_signal = None
        self._tweaked_normalized_signal = None
        self.tweak_spline = None
        self.strand = None
        self.fastq = None
        self.sequence_to_signal_mapping = None
        self.sequence = None

    @property
    def tweaked_normalized_signal(self):
        """read.py:94.  The batched estimator evaluates the tweak spline on the GPU for the aligned slice only and
        keeps the spline (``tweak_spline``); the whole-read array of the reference is produced on first access with
        the same scipy call."""
        if self._tweaked_normalized_signal is None and self.tweak_spline is not None:
            self._tweaked_normalized_signal = interpolate.splev(self.normalized_signal, self.tweak_spline)
        return self._tweaked_normalized_signal

    @tweaked_normalized_signal.setter
    def tweaked_normalized_signal(self, value):
        self._tweaked_normalized_signal = value

    @staticmethod
    def from_arrays(raw_signal, sequence, sequence_to_signal_mapping, name='read'):
        """Build a read from its three ingredients (read.py:10-17): raw samples, basecalled bases and the
        base index -> sample index map of the basecaller."""
        read = Read()
        read.raw_signal = numpy.asarray(raw_signal)
        read.sequence = numpy.array(list(sequence)) if isinstance(sequence, str) else numpy.asarray(sequence)
        read.sequence_to_signal_mapping = dict(sequence_to_signal_mapping)
        seq = ''.join(read.sequence)
        read.fastq = '@{}\n{}\n+\n{}\n'.format(name, seq, 'I' * len(seq))
        return read

    @staticmethod
    def load_from_fast5(filename, basecall_group, segmentation_group='Analyses/Segmentation_000'):
        raise NotImplementedError(
            'fast5 loading is host I/O outside the scope of nadavca_b200 (h5py is not available here); '
            'build reads with Read.from_arrays(raw_signal, sequence, sequence_to_signal_mapping)')

    @staticmethod
    def normalize_reads(reads, process_group=None):
        """One median / MAD pooled over all given reads, clipped to +-5 (read.py:67-81).

        With a torch.distributed `process_group` the reads of ALL ranks are pooled (every rank passes its shard): the
        two medians are exact order statistics found by bisection on the float64 bit pattern with one small
        all-reduce of counts per step, so a sharded job normalises exactly like the reference does on one host."""
        values = numpy.concatenate([numpy.asarray(read.raw_signal, dtype=float) for read in reads]) \
            if len(reads) else numpy.zeros(0)
        if process_group is None:
            shift = float(numpy.median(values))
            scale = float(numpy.median(abs(values - shift)))
        else:
            shift = distributed_median(values, process_group)
            scale = distributed_median(abs(values - shift), process_group)
        for read in reads:
            read.normalized_signal = numpy.clip((read.raw_signal - shift) / scale, -5, 5)

    def tweak_signal_normalization(self, alignment, expected_means, event_means=None):
        """Cubic smoothing spline from observed event means to expected levels (read.py:83-94).

        `event_means` (optional) are the per-event means of ``normalized_signal`` already computed on the device by
        ``dtw.Batch.event_means`` -- bit-identical to the ``numpy.mean`` calls of the reference's per-event Python
        loop, which at ~2000 events per read is the dominant host cost of ``estimate_snps``."""
        spline = self.fit_tweak_spline(alignment, expected_means, event_means)
        self.tweaked_normalized_signal = interpolate.splev(self.normalized_signal, spline)

    def fit_tweak_spline(self, alignment, expected_means, event_means=None):
        """The spline of ``tweak_signal_normalization`` (read.py:84-93) as FITPACK's (knots, coefficients, degree);
        also kept in ``self.tweak_spline``.  Pairs farther than 1 from their expected level are dropped, the rest is
        sorted by (mean, expected) like the reference's list of tuples."""
        if event_means is None:
            event_means = [numpy.mean(self.normalized_signal[event[0]: event[1]]) for event in alignment]
        means = numpy.asarray(event_means, dtype=float)
        expected = numpy.asarray(expected_means, dtype=float)
        with numpy.errstate(invalid='ignore'):
            keep = numpy.abs(expected - means) <= 1  # a NaN mean (empty event) fails the test, as in the reference
        means, expected = means[keep], expected[keep]
        order = numpy.lexsort((expected, means))
        means, expected = means[order], expected[order]
        self.tweak_spline = interpolate.splrep(means, expected, s=len(means))
        self._tweaked_normalized_signal = None
        return self.tweak_spline


def _ordered_keys(values):
    """float64 -> uint64 keys whose unsigned order is the numeric order of the floats (no NaNs expected)."""
    bits = numpy.ascontiguousarray(values, dtype=numpy.float64).view(numpy.uint64)
    sign = numpy.uint64(1) << numpy.uint64(63)
    return numpy.where(bits & sign, ~bits, bits | sign)


def _key_to_float(key):
    sign = numpy.uint64(1) << numpy.uint64(63)
    key = numpy.uint64(key)
    bits = (key & ~sign) if (key & sign) else ~key
    return float(numpy.array([bits], dtype=numpy.uint64).view(numpy.float64)[0])


def distributed_median(values, process_group):
    """numpy.median of the concatenation of every rank's `values` (mean of the two middle elements for an even
    count), computed without moving the samples: bisection over the 64-bit ordered keys, one all-reduce of a count
    per bit."""
    import torch
    import torch.distributed as dist
    keys = numpy.sort(_ordered_keys(values))
    device = 'cuda' if dist.get_backend(process_group) == 'nccl' else 'cpu'

    def total(x):
        t = torch.tensor([int(x)], dtype=torch.int64, device=device)
        dist.all_reduce(t, group=process_group)
        return int(t.item())

    n = total(len(keys))
    if n == 0:
        return float('nan')

    def kth(k):  # smallest key with at least k+1 pooled elements <= key
        lo, hi = 0, (1 << 64) - 1
        while lo < hi:
            mid = (lo + hi) >> 1
            if total(numpy.searchsorted(keys, numpy.uint64(mid), side='right')) >= k + 1:
                hi = mid
            else:
                lo = mid + 1
        return _key_to_float(lo)

    if n % 2:
        return kth(n // 2)
    return (kth(n // 2 - 1) + kth(n // 2)) / 2.0
