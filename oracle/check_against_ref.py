"""Pin the oracle: (1) the C port (oracle/nadavca_oracle.c) against the UNMODIFIED reference C++ compiled into
oracle/_ref by oracle/Makefile -- bit-identical doubles and ints on randomised cases; (2) the Python glue
restatement (oracle/oracle.py OracleEstimator) against the reference's own nadavca/estimator.py imported from
/root/reference under shims -- bit-identical chunks / posteriors / alignment tables.

Run here (needs /root/reference for part 2):  python oracle/check_against_ref.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def random_case(rng, k, cp, n, bw, mel, sparse=False, homopolymer=False):
    mean = rng.normal(0, 1.2, size=4 ** k)
    sigma = rng.uniform(0.2, 0.6, size=4 ** k)
    ref = rng.integers(0, 4, size=n)
    if homopolymer:
        ref[n // 3:n // 3 + k + 3] = ref[n // 3]
    lengths = np.maximum(mel, rng.poisson(6, size=n))
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]) + bw
    ids = np.zeros(n, dtype=int)
    padded = np.zeros(n + k, dtype=int)
    padded[cp:cp + n] = ref
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


def check_port_vs_ref(n_cases=60, seed=7):
    rng = np.random.default_rng(seed)
    bad = 0
    for case in range(n_cases):
        k = int(rng.integers(1, 5))
        cp = int(rng.integers(0, k))
        n = int(rng.integers(3, 40))
        bw = int(rng.integers(2, 25))
        mel = int(rng.integers(0, 4))
        mean, sigma, sig, ref, cb, ca, anc = random_case(rng, k, cp, n, bw, mel, sparse=case % 3 == 1,
                                                         homopolymer=case % 4 == 2)
        mp = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
        mr = orc.OracleModel(k, cp, 4, mean, sigma, 'ref')
        for flag in (False, True):
            a = orc.refine_alignment(sig, ref, cb, ca, anc, bw, mel, mp, flag)
            b = orc.refine_alignment(sig, ref, cb, ca, anc, bw, mel, mr, flag)
            if a != b:
                bad += 1
                print('refine mismatch', case, flag)
            a = np.array(orc.estimate_log_likelihoods(sig, ref, cb, ca, anc, bw, mel, mp, flag))
            b = np.array(orc.estimate_log_likelihoods(sig, ref, cb, ca, anc, bw, mel, mr, flag))
            if not np.array_equal(a, b):
                bad += 1
                print('ell mismatch', case, flag, np.nanmax(np.abs(a - b)))
        if mp.get_expected_signal(ref, cb, ca) != mr.get_expected_signal(ref, cb, ca):
            bad += 1
            print('expected_signal mismatch', case)
    print('port vs ref: %d cases, %d mismatches' % (n_cases, bad))
    return bad


def import_reference_package():
    """Import /root/reference/nadavca with the four shims of SURVEY.md 8(c)."""
    np.int = int
    np.float = float
    for name in ('h5py', 'simplesam'):
        sys.modules.setdefault(name, types.ModuleType(name))
    ref_dtw = orc.ref_module()
    import importlib.util
    pkg_dir = '/root/reference/nadavca'
    spec = importlib.util.spec_from_file_location('nadavca', os.path.join(pkg_dir, '__init__.py'),
                                                  submodule_search_locations=[pkg_dir])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules['nadavca'] = pkg
    sys.modules['nadavca.dtw'] = ref_dtw
    pkg.dtw = ref_dtw
    spec.loader.exec_module(pkg)
    return pkg


def check_glue_vs_reference(seed=3):
    if not os.path.isdir('/root/reference/nadavca'):
        print('glue check skipped: /root/reference absent')
        return 0
    import_reference_package()
    import nadavca.estimator as ref_est
    import nadavca.read as ref_read
    from nadavca_b200 import synthetic
    from nadavca_b200.kmer_model import load_kmer_model
    km = load_kmer_model(os.path.join(ROOT, 'nadavca_b200', 'default', 'kmer_model.hdf5'))
    cfg = dict(bandwidth=30, snp_prior_probability=0.001, min_event_length=2, model_wobbling=True,
               model_transitions=True, tweak_signal_normalization=True, normalization_event_length=10)
    genome = synthetic.make_genome(1500, seed=seed)
    bad = 0
    for tweak in (True, False):
        cfg['tweak_signal_normalization'] = tweak
        reads = [synthetic.make_read(genome, km, i, n_bases=70 + 5 * i, bandwidth=30, substitution_rate=0.03,
                                     jitter=6, start=(40 * i if i < 3 else 200 * i)) for i in range(6)]
        ref_reads = []
        for r in reads:  # the reference's Read class carries the same attributes
            rr = ref_read.Read()
            rr.raw_signal, rr.sequence, rr.sequence_to_signal_mapping = r.raw_signal, r.sequence, r.sequence_to_signal_mapping
            rr.truth = r.truth
            ref_reads.append(rr)
        ref_read.Read.normalize_reads(ref_reads)
        from nadavca_b200.read import Read
        Read.normalize_reads(reads)
        for a, b in zip(reads, ref_reads):
            if not np.array_equal(a.normalized_signal, b.normalized_signal):
                bad += 1
                print('normalize_reads mismatch')
        aligner = synthetic.SyntheticAligner(genome)
        ref_model = orc.ref_module().KmerModel(km.get_k(), km.get_central_position(), 4, km.mean.tolist(),
                                               km.sigma.tolist())
        ref_estimator = ref_est.ProbabilityEstimator(ref_model, aligner, cfg)
        om = orc.OracleModel(km.get_k(), km.get_central_position(), 4, km.mean, km.sigma, 'port')
        our = orc.OracleEstimator(om, aligner, cfg)
        for a, b in zip(reads, ref_reads):
            x, y = our.get_refined_alignment(a), ref_estimator.get_refined_alignment(b)
            if not np.array_equal(x[1], y[1]):
                bad += 1
                print('refined alignment table mismatch')
        xs = our.estimate_probabilities(genome, reads)
        ys = ref_estimator.estimate_probabilities(genome, ref_reads)
        if len(xs) != len(ys):
            bad += 1
            print('group count mismatch', len(xs), len(ys))
        for x, y in zip(xs, ys):
            if (x.start, x.end) != (y.start, y.end) or not np.array_equal(x.values, y.values) or \
                    not np.array_equal(x.coverage, y.coverage):
                bad += 1
                print('posterior chunk mismatch', x.start, y.start, np.abs(x.values - y.values).max())
        print('glue vs reference estimator.py (tweak=%s): %d groups compared' % (tweak, len(ys)))
    return bad


if __name__ == '__main__':
    if orc.ref_module() is None:
        print('oracle/_ref is not built and /root/reference is absent: nothing to check against')
        sys.exit(1)
    failures = check_port_vs_ref() + check_glue_vs_reference()
    print('FAILURES: %d' % failures)
    sys.exit(1 if failures else 0)
