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
    def normalize_reads(reads, process_group=None, device=None):
        """One median / MAD pooled over all given reads, clipped to +-5 (read.py:67-81).

        With a torch.distributed `process_group` the reads of ALL ranks are pooled (every rank passes its shard): the
        two medians are exact order statistics found by bisection on the float64 bit pattern with one small
        all-reduce of counts per step, so a sharded job normalises exactly like the reference does on one host.

        With `device` (a CUDA ordinal) the pooled samples are normalised ON THE GPU: exact radix select of the two
        medians (8 histogram passes each, the 256-bin histograms all-reduced over `process_group` when given) and the
        clip kernel; the result is bit-identical to the host path."""
        if device is not None:
            # the pooled samples go through two cached pinned staging buffers (H2D and D2H at PCIe speed instead of
            # pageable copies); every read gets its own copy of its slice
            normalized = _normalize_on_device([numpy.asarray(read.raw_signal, dtype=float) for read in reads],
                                              int(device), process_group)
            start = 0
            for read in reads:
                read.normalized_signal = normalized[start:start + len(read.raw_signal)].copy()
                start += len(read.raw_signal)
            return
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

    @staticmethod
    def normalize_each(reads, device):
        """``Read.normalize_reads([read])`` for every read on its own (what align_signal does, align_signal.py:54), as
        one batch on CUDA device `device`: exact per-read median / MAD by radix select, one CTA per read
        (csrc/select.cu); bit-identical to the host path."""
        import ctypes
        from . import _cabi
        if not reads:
            return
        lib = _cabi.require_device()
        raws = [numpy.asarray(read.raw_signal, dtype=float) for read in reads]
        off = numpy.zeros(len(reads) + 1, dtype=numpy.int64)
        off[1:] = numpy.cumsum([len(r) for r in raws])
        values = numpy.ascontiguousarray(numpy.concatenate(raws))
        out = numpy.empty_like(values)
        _cabi.check(lib.nvb_normalize_each(int(device), _cabi.ptr(values, ctypes.c_double), _cabi.ptr(off, ctypes.c_int64),
                                           len(reads), -5.0, 5.0, _cabi.ptr(out, ctypes.c_double), None,
                                           ctypes.c_void_p(0)), 'nvb_normalize_each')
        for i, read in enumerate(reads):
            read.normalized_signal = out[off[i]:off[i + 1]]

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
        self.tweak_spline = fit_spline(event_means, expected_means)
        self._tweaked_normalized_signal = None
        return self.tweak_spline


def fit_spline(event_means, expected_means):
    """read.py:87-93 for one read: (observed event mean, expected level) pairs -> scipy.interpolate.splrep."""
    means = numpy.asarray(event_means, dtype=float)
    expected = numpy.asarray(expected_means, dtype=float)
    with numpy.errstate(invalid='ignore'):
        keep = numpy.abs(expected - means) <= 1  # a NaN mean (empty event) fails the test, as in the reference
    means, expected = means[keep], expected[keep]
    order = numpy.lexsort((expected, means))
    means, expected = means[order], expected[order]
    return interpolate.splrep(means, expected, s=len(means))


class _Ready:
    """Result holder with the ``get()`` of the pool's result handle."""

    def __init__(self, value):
        self.value = value

    def get(self):
        return self.value


class _FitPool:
    """Persistent pool of ``nadavca_b200.fit_worker`` subprocesses.  Plain subprocesses talking pickle over pipes:
    unlike multiprocessing's spawn / forkserver workers they do not re-import the caller's ``__main__`` (an unguarded
    script would run again in every worker) and they never inherit this process's CUDA context."""

    def __init__(self, workers):
        import os
        import subprocess
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get('PYTHONPATH', ''),
                   OMP_NUM_THREADS='1', OPENBLAS_NUM_THREADS='1', MKL_NUM_THREADS='1')
        self.procs = [subprocess.Popen([sys.executable, '-m', 'nadavca_b200.fit_worker'], stdin=subprocess.PIPE,
                                       stdout=subprocess.PIPE, env=env) for _ in range(workers)]
        self.pending = []  # submitted, not yet collected: results come back through the pipes in this order

    def submit(self, jobs):
        """Deal the jobs to the workers in contiguous shares; returns a handle with ``get()``."""
        import pickle
        import struct
        n = len(self.procs)
        bounds = [len(jobs) * i // n for i in range(n + 1)]
        used = []
        for proc, lo, hi in zip(self.procs, bounds[:-1], bounds[1:]):
            if hi > lo:
                blob = pickle.dumps(jobs[lo:hi], protocol=pickle.HIGHEST_PROTOCOL)
                proc.stdin.write(struct.pack('<q', len(blob)))
                proc.stdin.write(blob)
                proc.stdin.flush()
                used.append(proc)
        handle = _PoolResult(self, used)
        self.pending.append(handle)
        return handle

    def close(self):
        for proc in self.procs:
            try:
                proc.stdin.close()
                proc.terminate()
            except Exception:
                pass


class _PoolResult:
    def __init__(self, pool, procs):
        self.pool, self.procs, self.value, self.error = pool, procs, None, None

    def get(self):
        while self.value is None and self.error is None:
            self.pool.pending.pop(0)._collect()  # earlier submissions first: the pipes are FIFO
        if self.error:
            raise RuntimeError('spline fit failed in a worker: {}'.format(self.error))
        return self.value

    def _collect(self):
        import pickle
        import struct
        if self.value is None:
            out = []
            error = None
            for proc in self.procs:
                head = proc.stdout.read(8)
                if len(head) < 8:
                    error = error or 'a spline-fit worker died'
                    continue
                status, payload = pickle.loads(proc.stdout.read(struct.unpack('<q', head)[0]))
                if status == 'ok':
                    out.extend(payload)
                else:
                    error = error or payload
            self.error = error
            self.value = out


_FIT_POOL = None


def fit_splines_async(jobs, workers=None):
    """``fit_spline`` for many reads; returns a handle whose ``get()`` gives the list of splines.  The FITPACK fit is
    host work that the reference does one read at a time (0.3-1 ms each, holding the GIL: threads do not help);
    batches of 32 reads or more go through a persistent pool of worker processes, which run the SAME scipy call, so
    the splines are identical.  NADAVCA_FIT_WORKERS=0 disables the pool."""
    global _FIT_POOL
    import os
    jobs = [(numpy.asarray(m, dtype=float), numpy.asarray(e, dtype=float)) for m, e in jobs]
    if workers is None:
        workers = int(os.environ.get('NADAVCA_FIT_WORKERS', min(32, max(1, (os.cpu_count() or 1) //
                                                                        int(os.environ.get('LOCAL_WORLD_SIZE', '1'))))))
    if workers <= 1 or len(jobs) < 32:
        return _Ready([fit_spline(*job) for job in jobs])
    if _FIT_POOL is None or _FIT_POOL[1] != workers:
        import atexit
        if _FIT_POOL is not None:
            _FIT_POOL[0].close()
        pool = _FitPool(workers)
        atexit.register(pool.close)
        _FIT_POOL = (pool, workers)
    return _FIT_POOL[0].submit(jobs)


_STAGING = {}


def _staging(name, count):
    """A cached pinned float64 host buffer of at least `count` elements (grow-only; `trim_staging()` frees them)."""
    import torch
    buf = _STAGING.get(name)
    if buf is None or buf.numel() < count:
        buf = torch.empty(max(int(count), 1), dtype=torch.float64).pin_memory()
        _STAGING[name] = buf
    return buf


def trim_staging():
    """Release the pinned staging buffers of the device normalisation."""
    _STAGING.clear()


def _normalize_on_device(values, device, process_group=None):
    """clip((values - median) / MAD, -5, 5) computed on CUDA device `device` (csrc/select.cu); `values` is this
    rank's share of the pooled samples: one float64 array or a list of them (concatenated straight into the pinned
    upload buffer).  The result is a view of a cached pinned buffer: copy what you keep."""
    import ctypes
    import torch
    from . import _cabi
    lib = _cabi.require_device()
    dev = torch.device('cuda', device)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    parts = values if isinstance(values, (list, tuple)) else [numpy.asarray(values, dtype=numpy.float64).reshape(-1)]
    n_local = int(sum(len(p) for p in parts))
    up = _staging('up', n_local)
    if n_local:
        numpy.concatenate(parts, out=up.numpy()[:n_local])
    d_values = up[:n_local].to(dev, non_blocking=True)
    dist = None
    if process_group is not None:
        import torch.distributed as dist
    count = torch.tensor([n_local], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(count, group=process_group)
    n = int(count.item())  # (synchronises: the upload buffer may be reused from here on)
    if n == 0:
        return numpy.zeros(0)

    def kth(k, absolute_deviation, shift):
        """k-th smallest (0-based) pooled value: most significant digit first, 8 bits per pass."""
        prefix = 0
        for fixed in range(0, 64, 8):
            hist = torch.zeros(256, dtype=torch.int64, device=dev)
            _cabi.check(lib.nvb_radix_histogram_d(device, ctypes.c_void_p(d_values.data_ptr()), n_local,
                                                  int(absolute_deviation), float(shift), prefix, fixed,
                                                  ctypes.c_void_p(hist.data_ptr()), sp), 'nvb_radix_histogram_d')
            if dist is not None:
                dist.all_reduce(hist, group=process_group)
            cum = numpy.cumsum(hist.cpu().numpy())
            digit = int(numpy.searchsorted(cum, k, side='right'))
            if digit:
                k -= int(cum[digit - 1])
            prefix = (prefix << 8) | digit
        return _key_to_float(prefix)

    def median(absolute_deviation, shift):
        if n % 2:
            return kth(n // 2, absolute_deviation, shift)
        return (kth(n // 2 - 1, absolute_deviation, shift) + kth(n // 2, absolute_deviation, shift)) / 2.0

    shift = median(False, 0.0)
    scale = median(True, shift)
    d_out = torch.empty_like(d_values)
    _cabi.check(lib.nvb_normalize_clip_d(device, ctypes.c_void_p(d_values.data_ptr()), n_local, shift, scale, -5.0, 5.0,
                                         ctypes.c_void_p(d_out.data_ptr()), sp), 'nvb_normalize_clip_d')
    down = _staging('down', n_local)
    down[:n_local].copy_(d_out, non_blocking=True)
    stream.synchronize()
    return down.numpy()[:n_local]


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
