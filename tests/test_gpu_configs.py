"""GPU parity at the scale of every BASELINE.json config, against the unmodified reference (oracle/_ref when it is
built, the bit-identical C port otherwise) run in a pool of worker processes on the box's host cores:

  configs[0]  100 reads (~2 kb) through align_signal: transitions, all three renormalisation rounds
  configs[3]  reads of ~100 kb bases / ~1 M samples: band width 150 (two reads) and 400 (one read), refine in both
              modes, estimate_log_likelihoods on one of them
  configs[4]  band widths 50 .. 1000 x both sweep schedules (where the rotating one is legal), refine in both modes
              and estimate

(configs[1] is the bench workload: bench.py checks its own benchmarked reads, see its `parity` block; configs[2] is
tests/test_gpu_consensus.py.)  Bars: event boundaries bit-exact, the only accepted deviation being a structural
tie as detected and counted by oracle/parity.py; raw log-likelihoods rtol 1e-9.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9


@pytest.fixture(scope='module')
def pool():
    import multiprocessing as mp
    p = mp.get_context('spawn').Pool(min(os.cpu_count() or 1, 48))
    yield p
    p.terminate()


def dp_args(read, aligner, bw, km):
    """What estimator.py:60-74,159-170 passes to the native module for one read."""
    from nadavca_b200.genome import Genome
    apx = aligner.get_signal_alignment(read, bw)
    s0, s1 = apx.signal_range
    a, b = apx.read_sequence_range
    k, cp = km.get_k(), km.get_central_position()
    return (read.normalized_signal[s0:s1], Genome.to_numerical(apx.reference_part),
            Genome.to_numerical(read.sequence[a - cp:a]), Genome.to_numerical(read.sequence[b:b + k - cp - 1]),
            apx.alignment)


def model_spec(km):
    from oracle import oracle as orc
    return (km.get_k(), km.get_central_position(), km.mean, km.sigma, 'ref' if orc.ref_module() is not None else 'port')


def oracle_model(km):
    from oracle import oracle as orc
    return orc.OracleModel(km.get_k(), km.get_central_position(), 4, km.mean, km.sigma, 'port')


class Tally:
    def __init__(self):
        self.exact = self.tie = 0

    def add(self, kind):
        if kind == 'tie':
            self.tie += 1
        else:
            self.exact += 1


def check_events(tally, ev, args, bw, mel, om, flag, want):
    from oracle import parity
    tally.add(parity.compare_events(ev, *args, bw, mel, om, flag, want=want))


# ---- configs[0] ------------------------------------------------------------------------------------------------

def test_config0_align_signal_100_reads(default_model, pool):
    """nadavca align: 100 synthetic reads (~2 kb bases, ~20 k samples) against a 50 kb reference, default config
    (bandwidth 150, min_event_length 2, transitions), three renormalisation rounds (align_signal.py:52-81)."""
    import nadavca_b200
    from nadavca_b200 import defaults, synthetic
    from nadavca_b200.genome import Genome
    from nadavca_b200.read import Read
    from oracle import parity
    from scipy.stats import linregress
    km = default_model
    cfg = dict(bandwidth=150, snp_prior_probability=0.001, min_event_length=2, model_wobbling=True,
               model_transitions=True, tweak_signal_normalization=True, normalization_event_length=10)
    genome = synthetic.make_genome(50_000, seed=0)
    aligner = synthetic.SyntheticAligner(genome)
    make = lambda: [synthetic.make_read(genome, km, 1000 + i) for i in range(100)]
    got = list(nadavca_b200.align_signal(None, make(), config=cfg, kmer_model=km, aligner=aligner, reference=genome,
                                         renorm_rounds=defaults.RENORM_ROUNDS))
    assert defaults.RENORM_ROUNDS == 3 and len(got) == 100

    # the same loop with the reference's refine_alignment on the host cores
    reads = make()
    spec, om = model_spec(km), oracle_model(km)
    for r in reads:
        Read.normalize_reads([r])

    def refine_all():
        args = [dp_args(r, aligner, 150, km) for r in reads]
        evs = pool.map(parity.job, [('refine', spec, a + (150, 2, True)) for a in args], chunksize=1)
        return args, evs

    def renormalise(evs):
        for r, ev in zip(reads, evs):
            apx = aligner.get_signal_alignment(r, 150)
            ev = np.asarray(ev) + apx.signal_range[0]
            expected = np.array(om.get_expected_signal(Genome.to_numerical(apx.reference_part), [], []))
            means = [np.mean(r.normalized_signal[s:e]) for s, e in ev]
            slope, intercept, _, _, _ = linregress(expected, means)
            r.normalized_signal = (r.normalized_signal - intercept) / slope

    _, evs = refine_all()          # first alignment
    renormalise(evs)               # round 0
    args, evs = refine_all()       # round 1
    tally = Tally()
    for (read_out, res), r, a, ev in zip(got, reads, args, evs):
        assert res is not None
        apx, table = res
        s0 = apx.signal_range[0]
        n = len(a[1])
        want_pos = (apx.reference_range[1] - 1 - np.arange(n)) if apx.reverse_complement \
            else (apx.reference_range[0] + np.arange(n))
        assert np.array_equal(table[:, 0], want_pos)
        check_events(tally, table[:, 1:] - s0, a, 150, 2, om, True, ev)
    renormalise(evs)               # round 2
    for (read_out, _), r in zip(got, reads):
        np.testing.assert_allclose(read_out.normalized_signal, r.normalized_signal, rtol=1e-12, atol=1e-12)
    print('configs[0]: %d alignments bit-exact, %d structural ties' % (tally.exact, tally.tie))
    assert tally.tie <= 2


# ---- configs[3] ------------------------------------------------------------------------------------------------

def test_config3_long_reads(default_model, pool):
    """Reads of ~100 kb bases / ~1 M samples: refine in both modes at band widths 150 and 400, and the SNP
    log-likelihoods of one read, against the reference."""
    from nadavca_b200 import dtw, synthetic
    from nadavca_b200.read import Read
    from oracle import parity
    km = default_model
    spec, om = model_spec(km), oracle_model(km)
    genome = synthetic.make_genome(400_000, seed=33)
    aligner = synthetic.SyntheticAligner(genome)
    plan = [(150, [synthetic.make_read(genome, km, 7000 + i, n_bases=100_000 + 777 * i) for i in range(2)]),
            (400, [synthetic.make_read(genome, km, 7100, n_bases=100_000, bandwidth=400)])]
    jobs, meta = [], []
    for bw, reads in plan:
        Read.normalize_reads(reads)
        for i, r in enumerate(reads):
            a = dp_args(r, aligner, bw, km)
            assert len(a[1]) >= 99_000 and len(a[0]) > 900_000
            for flag in (False, True):
                jobs.append(('refine', spec, a + (bw, 2, flag)))
                meta.append((bw, i, 'refine', flag))
            if bw == 150 and i == 0:
                jobs.append(('estimate', spec, a + (bw, 2, True)))
                meta.append((bw, i, 'estimate', True))
    # the estimate of a 100 kb read is ~1.3 G reference cells: start the oracle first, run the GPU meanwhile
    order = sorted(range(len(jobs)), key=lambda j: meta[j][2] != 'estimate')
    pending = {j: pool.apply_async(parity.job, (jobs[j],)) for j in order}
    tally, worst = Tally(), 0.0
    for bw, reads in plan:
        args = [dp_args(r, aligner, bw, km) for r in reads]
        with dtw.Batch(km, *[list(x) for x in zip(*args)], bw, 2) as batch:
            got = {}
            for flag in (False, True):
                batch.refine(flag)
                events, status = batch.events()
                assert np.all(status == 0)
                got[flag] = [e.copy() for e in events]
            lls = None
            if bw == 150:
                batch.estimate(True)
                lls, _ = batch.log_likelihoods()
                lls = [x.copy() for x in lls]
        for j, (mbw, i, kind, flag) in enumerate(meta):
            if mbw != bw:
                continue
            want = pending[j].get(timeout=3000)
            if kind == 'refine':
                check_events(tally, got[flag][i], args[i], bw, 2, om, flag, want)
            else:
                worst = max(worst, parity.ll_max_rel(lls[i], want))
                np.testing.assert_allclose(lls[i], want, rtol=LL_RTOL)
    print('configs[3]: %d alignments of ~100 kb reads bit-exact, %d structural ties, max rel LL error %.2e'
          % (tally.exact, tally.tie, worst))
    assert tally.exact + tally.tie == 6


# ---- configs[4] ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize('bw', [50, 150, 400, 650, 1000])
def test_config4_band_widths(default_model, pool, bw):
    """Band-width sweep 50 .. 1000 on ~2 kb reads: refine in both modes through both sweep schedules (the rotating
    one exists for band rows up to 640 columns) and estimate_log_likelihoods, against the reference."""
    from nadavca_b200 import dtw, synthetic
    from nadavca_b200.read import Read
    from oracle import parity
    km = default_model
    spec, om = model_spec(km), oracle_model(km)
    genome = synthetic.make_genome(100_000, seed=44)
    aligner = synthetic.SyntheticAligner(genome)
    reads = [synthetic.make_read(genome, km, 8000 + 10 * bw + i, n_bases=1700 + 300 * i, bandwidth=bw) for i in range(3)]
    Read.normalize_reads(reads)
    args = [dp_args(r, aligner, bw, km) for r in reads]
    jobs = [('refine', spec, a + (bw, 2, flag)) for a in args for flag in (False, True)]
    jobs.append(('estimate', spec, args[0] + (bw, 2, True)))
    if bw <= 150:
        jobs.append(('estimate', spec, args[1] + (bw, 2, False)))
    pending = [pool.apply_async(parity.job, (j,)) for j in jobs]
    results = {}
    schedules = ['s', 'r'] if 2 * bw + 1 <= 640 else ['s']
    try:
        for sched in schedules:
            dtw.set_sweep_schedule(sched)
            with dtw.Batch(km, *[list(x) for x in zip(*args)], bw, 2) as batch:
                assert max(int((be - bs + 1).max()) for bs, be in batch.bands()) >= 2 * bw + 1
                for flag in (False, True):
                    batch.refine(flag)
                    events, status = batch.events()
                    assert np.all(status == 0)
                    results[(sched, 'refine', flag)] = [e.copy() for e in events]
                batch.estimate(True)
                results[(sched, 'wobble')] = batch.log_likelihoods()[0][0].copy()
                if bw <= 150:
                    batch.estimate(False)
                    results[(sched, 'plain')] = batch.log_likelihoods()[0][1].copy()
    finally:
        dtw.set_sweep_schedule(None)
    want = [p.get(timeout=3000) for p in pending]
    tally, worst = Tally(), 0.0
    for sched in schedules:
        j = 0
        for i, a in enumerate(args):
            for flag in (False, True):
                check_events(tally, results[(sched, 'refine', flag)][i], a, bw, 2, om, flag, want[j])
                j += 1
        worst = max(worst, parity.ll_max_rel(results[(sched, 'wobble')], want[j]))
        np.testing.assert_allclose(results[(sched, 'wobble')], want[j], rtol=LL_RTOL)
        if bw <= 150:
            worst = max(worst, parity.ll_max_rel(results[(sched, 'plain')], want[j + 1]))
            np.testing.assert_allclose(results[(sched, 'plain')], want[j + 1], rtol=LL_RTOL)
    print('configs[4] bandwidth %d, schedules %s: %d alignments bit-exact, %d structural ties, max rel LL error %.2e'
          % (bw, '+'.join(schedules), tally.exact, tally.tie, worst))
    assert tally.tie <= 1
