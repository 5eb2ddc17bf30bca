import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def lib_built():
    from nadavca_b200 import build
    return build.build()


@pytest.fixture(scope='session')
def default_model(lib_built):
    from nadavca_b200.kmer_model import KmerModel
    return KmerModel.load_from_hdf5(os.path.join(ROOT, 'nadavca_b200', 'default', 'kmer_model.hdf5'))


def make_case(rng, k, cp, n, bw, mel, sparse=False, homopolymer=False, spacing=6):
    """Random model + one synthetic read slice: (mean, sigma, signal, reference, ctx_before, ctx_after, anchors)."""
    mean = rng.normal(0, 1.2, size=4 ** k)
    sigma = rng.uniform(0.2, 0.6, size=4 ** k)
    ref = rng.integers(0, 4, size=n)
    if homopolymer:
        ref[n // 3:n // 3 + k + 3] = ref[n // 3]
    lengths = np.maximum(mel, rng.poisson(spacing, size=n))
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]) + bw
    padded = np.zeros(n + k, dtype=int)
    padded[cp:cp + n] = ref
    ids = np.zeros(n, dtype=int)
    for j in range(k):
        ids = ids * 4 + padded[j:j + n]
    sig = np.concatenate([rng.normal(0, 1, bw), np.repeat(mean[ids], lengths) + rng.normal(0, 0.4, lengths.sum()),
                          rng.normal(0, 1, bw)])
    sig = np.clip(sig, -5, 5)
    anchors = np.stack([np.clip(starts + rng.integers(-bw // 3 - 1, bw // 3 + 2, size=n), 0, len(sig) - 1),
                        np.arange(n)], axis=1)
    anchors[:, 0] = np.maximum.accumulate(anchors[:, 0])
    if sparse:
        keep = np.sort(rng.choice(n, size=min(n, max(2, n // 4)), replace=False))
        keep[0], keep[-1] = 0, n - 1
        anchors = anchors[np.unique(keep)]
    cb = rng.integers(0, 4, size=rng.integers(0, cp + 1))
    ca = rng.integers(0, 4, size=rng.integers(0, k - cp))
    return mean, sigma, sig, ref, cb, ca, anchors


GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


class Golden:
    """Flat npz written by oracle/make_golden.py; keys look like 'case/field'."""

    def __init__(self, name):
        self.data = np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)

    def __getitem__(self, key):
        return self.data[key]

    def case(self, tag):
        pre = tag + '/'
        return {k[len(pre):]: self.data[k] for k in self.data.files if k.startswith(pre)}


@pytest.fixture(scope='session')
def golden_dp():
    return Golden('dp_cases.npz')


@pytest.fixture(scope='session')
def golden_estimator():
    return Golden('estimator_cases.npz')


def golden_reads(golden):
    """Rebuild the stored synthetic reads of estimator_cases.npz (raw signal, sequence, mapping, truth)."""
    from nadavca_b200.read import Read
    reads = []
    for i in range(int(golden['n_reads'])):
        seq = golden['read%d/sequence' % i]
        mapping = golden['read%d/mapping' % i]
        read = Read.from_arrays(golden['read%d/raw_signal' % i], seq, {b: int(s) for b, s in enumerate(mapping)},
                                name='golden_%d' % i)
        start, n, reverse, flank = (int(x) for x in golden['read%d/truth' % i])
        read.truth = {'start': start, 'n': n, 'reverse': bool(reverse), 'flank': flank}
        reads.append(read)
    return reads


GOLDEN_CONFIG = dict(bandwidth=30, snp_prior_probability=0.001, min_event_length=2, model_wobbling=True,
                     model_transitions=True, tweak_signal_normalization=True, normalization_event_length=10)


@pytest.fixture(scope='session')
def default_model_host():
    """The shipped 6-mer model described on the host (no CUDA needed until its device handle is used)."""
    from nadavca_b200.kmer_model import KmerModel
    return KmerModel.load_from_hdf5(os.path.join(ROOT, 'nadavca_b200', 'default', 'kmer_model.hdf5'))
