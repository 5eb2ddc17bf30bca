"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.   *** TEST INFRASTRUCTURE ***

The reference has no tests, fixtures or known-answer vectors of its own, so the vectors are
outputs of the reference itself, run in the build container:

  * dp_cases.npz        -- `nadavca.dtw` (reference nadavca/dtw/*.cpp compiled unmodified into oracle/_ref by
                           oracle/Makefile): refine_alignment (both model_transitions), estimate_log_likelihoods
                           (both model_wobbling) and KmerModel.get_expected_signal on seeded random cases
                           (k 1..4, min_event_length 0..3, bandwidth 2..24, sparse anchors, homopolymer runs,
                           infeasible bands), on the SURVEY.md section 4 toy vectors and on two reads simulated from
                           the shipped 6-mer model.
  * estimator_cases.npz -- the reference's own Python glue (/root/reference/nadavca/estimator.py and read.py,
                           imported under the four shims of SURVEY.md 8c) on top of that module:
                           ProbabilityEstimator.get_refined_alignment, ._estimate_log_likelihoods and
                           .estimate_probabilities on seeded synthetic reads (both strands, overlapping and disjoint
                           chunks, with and without the spline tweak) and calculate_meth_scores of
                           nadavca/detect_meth.py on those alignments.  The reads themselves (raw signal, sequence,
                           base->sample map, truth) are stored so that the tests do not depend on the generator.

Needs /root/reference (absent on the GPU box): run here with  `python oracle/make_golden.py`  and commit the .npz.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from oracle.check_against_ref import import_reference_package, random_case  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

TOY_SIGNAL = [0.1, -0.1, 0.2, 1.1, 0.9, 1.0, 2.1, 1.9, 2.2, 3.0, 3.1, 2.9]
TOY_ANCHORS = [[0, 0], [3, 1], [6, 2], [9, 3]]


def run_ref_case(out, tag, k, cp, mean, sigma, sig, ref, cb, ca, anc, bw, mel):
    model = orc.OracleModel(k, cp, 4, mean, sigma, 'ref')
    out[tag + '/params'] = np.array([k, cp, bw, mel], dtype=np.int64)
    out[tag + '/mean'] = np.asarray(mean, dtype=np.float64)
    out[tag + '/sigma'] = np.asarray(sigma, dtype=np.float64)
    out[tag + '/signal'] = np.asarray(sig, dtype=np.float64)
    out[tag + '/reference'] = np.asarray(ref, dtype=np.int32)
    out[tag + '/context_before'] = np.asarray(cb, dtype=np.int32)
    out[tag + '/context_after'] = np.asarray(ca, dtype=np.int32)
    out[tag + '/anchors'] = np.asarray(anc, dtype=np.int32).reshape(-1, 2)
    for flag in (0, 1):
        ev = orc.refine_alignment(sig, ref, cb, ca, anc, bw, mel, model, bool(flag))
        out[tag + '/events%d' % flag] = np.asarray(ev, dtype=np.int32).reshape(-1, 2)
        ll = orc.estimate_log_likelihoods(sig, ref, cb, ca, anc, bw, mel, model, bool(flag))
        out[tag + '/ll%d' % flag] = np.asarray(ll, dtype=np.float64)
    out[tag + '/expected'] = np.asarray(model.get_expected_signal(ref, cb, ca), dtype=np.float64)


def make_dp_cases():
    out = {}
    names = []
    # SURVEY.md section 4 toy vectors
    toy = dict(k=1, cp=0, mean=[0., 1., 2., 3.], sigma=[.5] * 4)
    run_ref_case(out, 'toy0', sig=TOY_SIGNAL, ref=[0, 1, 2, 3], cb=[], ca=[], anc=TOY_ANCHORS, bw=3, mel=2, **toy)
    run_ref_case(out, 'toy1', sig=TOY_SIGNAL, ref=[0, 0, 1, 1], cb=[], ca=[], anc=TOY_ANCHORS, bw=3, mel=2, **toy)
    run_ref_case(out, 'toy2', sig=TOY_SIGNAL[:3], ref=[0, 1, 2, 3], cb=[], ca=[], anc=[[0, 0]], bw=3, mel=2, **toy)
    names += ['toy0', 'toy1', 'toy2']
    rng = np.random.default_rng(20261018)
    for case in range(48):
        k = int(rng.integers(1, 5))
        cp = int(rng.integers(0, k))
        n = int(rng.integers(1, 48)) if case else 1
        bw = int(rng.integers(2, 25))
        mel = int(case % 4)
        mean, sigma, sig, ref, cb, ca, anc = random_case(rng, k, cp, n, bw, mel, sparse=case % 3 == 1,
                                                         homopolymer=case % 4 == 2)
        tag = 'rnd%02d' % case
        run_ref_case(out, tag, k, cp, mean, sigma, sig, ref, cb, ca, anc, bw, mel)
        names.append(tag)
    # an infeasible band in a larger model: 10 bases, 12 samples, min_event_length 3
    rng = np.random.default_rng(5)
    mean = rng.normal(0, 1, size=16)
    run_ref_case(out, 'nopath', 2, 1, mean, np.full(16, 0.4), rng.normal(0, 1, 12), rng.integers(0, 4, 10), [], [],
                 [[0, 0], [11, 9]], 2, 3)
    names.append('nopath')
    # shipped 6-mer model, both strands, bandwidth 30
    from nadavca_b200 import synthetic
    from nadavca_b200.kmer_model import load_kmer_model
    from nadavca_b200.read import Read
    km = load_kmer_model(os.path.join(ROOT, 'nadavca_b200', 'default', 'kmer_model.hdf5'))
    genome = synthetic.make_genome(2500, seed=2)
    reads = [synthetic.make_read(genome, km, 40 + i, n_bases=110, bandwidth=30, strand=s, substitution_rate=0.03,
                                 jitter=6) for i, s in enumerate('+-')]
    Read.normalize_reads(reads)
    aligner = synthetic.SyntheticAligner(genome)
    for i, r in enumerate(reads):
        apx = aligner.get_signal_alignment(r, 30)
        s0, s1 = apx.signal_range
        a, b = apx.read_sequence_range
        tag = 'model6_%d' % i
        run_ref_case(out, tag, 6, 2, km.mean, km.sigma, r.normalized_signal[s0:s1],
                     orc.to_numerical(apx.reference_part), orc.to_numerical(r.sequence[a - 2:a]),
                     orc.to_numerical(r.sequence[b:b + 3]), apx.alignment, 30, 2)
        names.append(tag)
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(GOLDEN, 'dp_cases.npz'), **out)
    print('dp_cases.npz: %d cases' % len(names))


def make_estimator_cases():
    import_reference_package()
    import nadavca.estimator as ref_est
    import nadavca.read as ref_read
    from nadavca_b200 import synthetic
    from nadavca_b200.kmer_model import load_kmer_model
    km = load_kmer_model(os.path.join(ROOT, 'nadavca_b200', 'default', 'kmer_model.hdf5'))
    ref_model = orc.ref_module().KmerModel(km.get_k(), km.get_central_position(), 4, km.mean.tolist(),
                                           km.sigma.tolist())
    out = {}
    genome = synthetic.make_genome(1800, seed=11)
    out['genome'] = genome
    n_reads = 7
    starts = [30, 70, 110, 600, 640, 1200, 1290]   # two overlap groups of 3 and 2, a touching pair (Q9), a single
    lens = [90, 80, 100, 85, 95, 90, 70]           # 1200+90 == 1290: the '>=' rule opens a new group
    strands = '+-+-+-+'
    reads = [synthetic.make_read(genome, km, 70 + i, n_bases=lens[i], bandwidth=30, substitution_rate=0.03, jitter=6,
                                 start=starts[i], strand=strands[i]) for i in range(n_reads)]
    out['n_reads'] = np.array(n_reads)
    for i, r in enumerate(reads):
        out['read%d/raw_signal' % i] = np.asarray(r.raw_signal, dtype=np.float64)
        out['read%d/sequence' % i] = np.asarray(r.sequence)
        out['read%d/mapping' % i] = np.array([r.sequence_to_signal_mapping[b] for b in range(len(r.sequence))],
                                             dtype=np.int64)
        t = r.truth
        out['read%d/truth' % i] = np.array([t['start'], t['n'], int(t['reverse']), t['flank']], dtype=np.int64)
    aligner = synthetic.SyntheticAligner(genome)
    for tweak in (1, 0):
        cfg = dict(bandwidth=30, snp_prior_probability=0.001, min_event_length=2, model_wobbling=True,
                   model_transitions=True, tweak_signal_normalization=bool(tweak), normalization_event_length=10)
        ref_reads = []
        for r in reads:
            rr = ref_read.Read()
            rr.raw_signal, rr.sequence = r.raw_signal, r.sequence
            rr.sequence_to_signal_mapping = r.sequence_to_signal_mapping
            rr.truth = r.truth
            ref_reads.append(rr)
        ref_read.Read.normalize_reads(ref_reads)
        est = ref_est.ProbabilityEstimator(ref_model, aligner, cfg)
        pre = 'tweak%d/' % tweak
        for i, rr in enumerate(ref_reads):
            if tweak:
                out['read%d/normalized_signal' % i] = np.asarray(rr.normalized_signal, dtype=np.float64)
            apx, table = est.get_refined_alignment(rr)
            out[pre + 'read%d/alignment_table' % i] = np.asarray(table, dtype=np.int64)
            chunk = est._estimate_log_likelihoods(genome, rr)
            out[pre + 'read%d/chunk_range' % i] = np.array([chunk.start, chunk.end], dtype=np.int64)
            out[pre + 'read%d/chunk_values' % i] = np.asarray(chunk.values, dtype=np.float64)
            ind = est.estimate_probabilities(genome, [rr])[0]
            out[pre + 'read%d/independent_probabilities' % i] = np.asarray(ind.values, dtype=np.float64)
        if tweak:
            # nadavca/detect_meth.py:26-55 on the reference's own refined alignments (pattern "CG")
            import nadavca.detect_meth as ref_meth
            for i, rr in enumerate(ref_reads):
                apx, table = est.get_refined_alignment(rr)
                cut = rr.normalized_signal[table[0][1]:table[-1][2]]
                feats = ref_meth.calculate_meth_scores(cut, table, apx, 'CG', ref_model)
                out['meth/read%d/positions' % i] = np.array([f[0] for f in feats], dtype=np.int64)
                out['meth/read%d/contexts' % i] = np.array([f[1] for f in feats], dtype='U11')
                out['meth/read%d/scores' % i] = np.array([f[2] for f in feats], dtype=np.float64).reshape(-1, 11)
                out['meth/read%d/aggregated' % i] = np.array([ref_meth.maxs3(f[2]) for f in feats], dtype=np.float64)
        groups = est.estimate_probabilities(genome, ref_reads)
        out[pre + 'n_groups'] = np.array(len(groups))
        for g, chunk in enumerate(groups):
            out[pre + 'group%d/range' % g] = np.array([chunk.start, chunk.end], dtype=np.int64)
            out[pre + 'group%d/probabilities' % g] = np.asarray(chunk.values, dtype=np.float64)
            out[pre + 'group%d/coverage' % g] = np.asarray(chunk.coverage, dtype=np.int64)
        print('estimator_cases (tweak=%d): %d reads, %d consensus groups' % (tweak, n_reads, len(groups)))
    np.savez_compressed(os.path.join(GOLDEN, 'estimator_cases.npz'), **out)


if __name__ == '__main__':
    if orc.ref_module() is None or not os.path.isdir('/root/reference/nadavca'):
        sys.exit('make_golden.py needs /root/reference and oracle/_ref (make -C oracle ref)')
    os.makedirs(GOLDEN, exist_ok=True)
    make_dp_cases()
    make_estimator_cases()
