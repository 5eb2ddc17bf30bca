"""GPU parity tests proper: the CUDA path (through the C ABI / Python drop-in) against the CPU oracle on the same
seeded inputs.  Bars: alignment events and bands bit-exact (ints); log-likelihoods within LL_RTOL relative
(north_star allows 1e-5; the fp64 kernels are held to 1e-9)."""
import numpy as np
import pytest

from conftest import make_case

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9
LL_ATOL = 1e-9

SIG = [0.1, -0.1, 0.2, 1.1, 0.9, 1.0, 2.1, 1.9, 2.2, 3.0, 3.1, 2.9]
ANC = [[0, 0], [3, 1], [6, 2], [9, 3]]


@pytest.fixture(scope='module')
def toy(lib_built):
    from nadavca_b200.dtw import KmerModel
    return KmerModel(1, 0, 4, [0, 1, 2, 3], [.5] * 4)


def test_known_answers_refine(toy):
    """Known-answer vectors produced by the compiled reference."""
    from nadavca_b200 import dtw
    for flag in (True, False):
        out = dtw.refine_alignment(signal=SIG, reference=[0, 1, 2, 3], context_before=[], context_after=[],
                                   approximate_alignment=ANC, bandwidth=3, min_event_length=2, kmer_model=toy,
                                   model_transitions=flag)
        assert out == [[0, 3], [3, 6], [6, 9], [9, 11]]
    out = dtw.refine_alignment(signal=SIG, reference=[0, 0, 1, 1], context_before=[], context_after=[],
                               approximate_alignment=ANC, bandwidth=3, min_event_length=2, kmer_model=toy,
                               model_transitions=True)
    assert out == [[0, 2], [2, 4], [4, 6], [6, 8]]
    out = dtw.refine_alignment(signal=SIG[:3], reference=[0, 1, 2, 3], context_before=[], context_after=[],
                               approximate_alignment=[[0, 0]], bandwidth=3, min_event_length=2, kmer_model=toy,
                               model_transitions=True)
    assert out == []


def test_known_answers_log_likelihoods(toy):
    from nadavca_b200 import dtw
    ll = dtw.estimate_log_likelihoods(signal=SIG, reference=[0, 1, 2, 3], context_before=[], context_after=[],
                                      approximate_alignment=ANC, bandwidth=3, min_event_length=2, kmer_model=toy,
                                      model_wobbling=True)
    np.testing.assert_allclose(ll[0], [-0.013043763389, -4.261977157088, -16.055214132595, -35.692998888034],
                               rtol=1e-10)
    np.testing.assert_allclose(ll[3], [-20.042425824037, -7.636074636067, -2.296276451577, -0.013043763389],
                               rtol=1e-10)
    ll = dtw.estimate_log_likelihoods(signal=SIG, reference=[0, 1, 2, 3], context_before=[], context_after=[],
                                      approximate_alignment=ANC, bandwidth=3, min_event_length=2, kmer_model=toy,
                                      model_wobbling=False)
    np.testing.assert_allclose(ll[0], [-0.500222199981, -1.982921519971, -7.528306532021, -19.725561807744],
                               rtol=1e-10)


TIE_LOG = {'exact': 0, 'tie': 0}  # alignments compared so far in this process / accepted as structural ties


def assert_same_path(ev, case, bw, mel, om, flag):
    """Events must equal the oracle's bit for bit.  The one exception is detected per read, not assumed per batch
    (oracle/parity.py): rows next to a neighbour with an IDENTICAL emission tie mathematically and the reference's own
    argmax there is rounding noise; a path that differs ONLY at such rows and has the same max-product score under
    the oracle's posteriors is counted as a tie."""
    from oracle import parity
    kind = parity.compare_events(ev, case[2], case[3], case[4], case[5], case[6], bw, mel, om, flag)
    TIE_LOG[kind] += 1
    return kind == 'exact'


def _compare_batch(rng, cases, k, cp, mel, bw, mean, sigma):
    """Run a ragged batch through the GPU and every read through the oracle."""
    from nadavca_b200 import dtw
    from oracle import oracle as orc
    gm = dtw.KmerModel(k, cp, 4, mean, sigma)
    om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
    sigs = [c[2] for c in cases]
    refs = [c[3] for c in cases]
    cbs = [c[4] for c in cases]
    cas = [c[5] for c in cases]
    ancs = [c[6] for c in cases]
    with dtw.Batch(gm, sigs, refs, cbs, cas, ancs, bw, mel) as batch:
        bands = batch.bands()
        for (bs, be), c in zip(bands, cases):
            obs, obe = orc.band_bounds(c[6], len(c[2]), len(c[3]), bw)
            assert np.array_equal(bs, obs) and np.array_equal(be, obe)
        for flag in (False, True):
            batch.refine(flag)
            events, status = batch.events()
            for ci, (ev, st, c) in enumerate(zip(events, status, cases)):
                assert_same_path(ev, c, bw, mel, om, flag)
                assert (ev is None) == (st == 1)
            batch.estimate(flag)
            lls, status = batch.log_likelihoods()
            for ll, c in zip(lls, cases):
                want = np.array(orc.estimate_log_likelihoods(c[2], c[3], c[4], c[5], c[6], bw, mel, om, flag))
                finite = np.isfinite(want)
                assert np.array_equal(np.isfinite(ll), finite)
                np.testing.assert_allclose(ll[finite], want[finite], rtol=LL_RTOL, atol=LL_ATOL)
        exp = gm.get_expected_signal_batch(refs, cbs, cas)
        for e, c in zip(exp, cases):
            assert e.tolist() == om.get_expected_signal(c[3], c[4], c[5])


@pytest.mark.parametrize('mel', [0, 1, 2, 3])
@pytest.mark.parametrize('k,cp', [(1, 0), (3, 1), (4, 2), (6, 2)])
def test_random_batches_match_oracle(lib_built, k, cp, mel):
    rng = np.random.default_rng(100 * k + mel)
    bw = int(rng.integers(3, 20))
    mean = rng.normal(0, 1.2, size=4 ** k)
    sigma = rng.uniform(0.2, 0.6, size=4 ** k)
    cases = []
    for i in range(12):
        n = int(rng.integers(1, 90)) if i else 1
        c = make_case(rng, k, cp, n, bw, mel, sparse=i % 3 == 1, homopolymer=i % 4 == 2)
        cases.append((mean, sigma) + c[2:])
    # cases i % 4 == 2 carry a homopolymer run longer than k: the structural ties of oracle/parity.py
    before = dict(TIE_LOG)
    _compare_batch(rng, cases, k, cp, mel, bw, mean, sigma)
    print('k=%d mel=%d: %d alignments exact, %d structural ties' % (k, mel, TIE_LOG['exact'] - before['exact'],
                                                                    TIE_LOG['tie'] - before['tie']))


def test_no_path_and_mixed_status(lib_built):
    """A read whose band cannot hold min_event_length samples per base has no path (reference returns [])."""
    rng = np.random.default_rng(5)
    k, cp, mel, bw = 2, 1, 3, 2
    mean = rng.normal(0, 1, size=16)
    sigma = np.full(16, 0.4)
    good = make_case(rng, k, cp, 20, bw, mel, spacing=8)
    sig = rng.normal(0, 1, 12)
    bad = (mean, sigma, sig, rng.integers(0, 4, 10), [], [], np.array([[0, 0], [11, 9]]))
    cases = [(mean, sigma) + good[2:], bad, (mean, sigma) + make_case(rng, k, cp, 33, bw, mel, spacing=8)[2:]]
    _compare_batch(rng, cases, k, cp, mel, bw, mean, sigma)


def test_default_model_read_matches_oracle(default_model):
    """Shipped 6-mer model, default bandwidth 150 / min_event_length 2, one ~300-base synthetic read per strand."""
    from nadavca_b200 import dtw, synthetic
    from nadavca_b200.read import Read
    from oracle import oracle as orc
    km = default_model
    om = orc.OracleModel(km.get_k(), km.get_central_position(), 4, km.mean, km.sigma, 'port')
    genome = synthetic.make_genome(3000, seed=1)
    reads = [synthetic.make_read(genome, km, i, n_bases=300, strand=s, substitution_rate=0.02)
             for i, s in enumerate('+-')]
    Read.normalize_reads(reads)
    aligner = synthetic.SyntheticAligner(genome)
    args = []
    for r in reads:
        apx = aligner.get_signal_alignment(r, 150)
        s0, s1 = apx.signal_range
        ref = orc.to_numerical(apx.reference_part)
        a, b = apx.read_sequence_range
        cb = orc.to_numerical(r.sequence[a - 2:a])
        ca = orc.to_numerical(r.sequence[b:b + 3])
        args.append((r.normalized_signal[s0:s1], ref, cb, ca, apx.alignment))
    with dtw.Batch(km, *[list(x) for x in zip(*args)], 150, 2) as batch:
        for flag in (True, False):
            batch.refine(flag)
            events, _ = batch.events()
            for ev, a in zip(events, args):
                assert ev.tolist() == orc.refine_alignment(*a, 150, 2, om, flag)
        batch.estimate(True)
        lls, _ = batch.log_likelihoods()
        for ll, a in zip(lls, args):
            want = np.array(orc.estimate_log_likelihoods(*a, 150, 2, om, True))
            np.testing.assert_allclose(ll, want, rtol=LL_RTOL)
        counts = batch.cell_counts(True)
        want = {key: 0 for key in counts}
        for a in args:
            c = orc.count_cells(a[4], len(a[0]), len(a[1]), 150, 6, 2)
            for key in want:
                want[key] += c[key]
        assert counts == want


def test_workspace_waves_give_identical_results(default_model):
    """Forcing the batch through several waves (small workspace limit) must not change any result."""
    from nadavca_b200 import dtw
    rng = np.random.default_rng(11)
    k, cp, mel, bw = 3, 1, 2, 10
    mean = rng.normal(0, 1.2, size=64)
    sigma = rng.uniform(0.2, 0.6, size=64)
    cases = [make_case(rng, k, cp, int(rng.integers(20, 60)), bw, mel) for _ in range(9)]
    gm = dtw.KmerModel(k, cp, 4, mean, sigma)
    lists = [[c[i] for c in cases] for i in (2, 3, 4, 5, 6)]
    out = []
    for limit in (0, 90_000):
        with dtw.Batch(gm, *lists, bw, mel, workspace_limit=limit) as batch:
            batch.refine(True)
            ev, _ = batch.events()
            batch.estimate(True)
            ll, _ = batch.log_likelihoods()
            out.append(([e.copy() for e in ev], [x.copy() for x in ll]))
    for a, b in zip(out[0][0], out[1][0]):
        assert np.array_equal(a, b)
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)


def test_batches_of_one_model_share_the_workspace(lib_built):
    """The DP workspace belongs to the model: interleaved runs on two batches must not disturb each other's resident
    results, and nvb_trim_memory must leave live batches intact."""
    from nadavca_b200 import _cabi, dtw
    rng = np.random.default_rng(21)
    k, cp, mel, bw = 3, 1, 2, 9
    mean = rng.normal(0, 1.2, size=64)
    sigma = rng.uniform(0.2, 0.6, size=64)
    gm = dtw.KmerModel(k, cp, 4, mean, sigma)

    def lists(cases):
        return [[c[i] for c in cases] for i in (2, 3, 4, 5, 6)]
    small = [make_case(rng, k, cp, int(rng.integers(10, 30)), bw, mel) for _ in range(3)]
    big = [make_case(rng, k, cp, int(rng.integers(60, 90)), bw, mel) for _ in range(5)]
    with dtw.Batch(gm, *lists(small), bw, mel) as a, dtw.Batch(gm, *lists(big), bw, mel) as b:
        a.refine(True)
        ev_a = [e.copy() for e in a.events()[0]]
        b.estimate(True)                      # grows and overwrites the shared workspace
        ll_b = [x.copy() for x in b.log_likelihoods()[0]]
        a.estimate(True)
        ll_a = [x.copy() for x in a.log_likelihoods()[0]]
        b.refine(True)
        assert _cabi.load().nvb_trim_memory(gm.device) == 0
        for x, y in zip(a.events()[0], ev_a):          # resident results of `a` survived b's runs and the trim
            assert np.array_equal(x, y)
        for x, y in zip(b.log_likelihoods()[0], ll_b):
            assert np.array_equal(x, y)
    # the same work on fresh, separate batches gives identical numbers
    with dtw.Batch(gm, *lists(small), bw, mel) as a2:
        a2.estimate(True)
        for x, y in zip(a2.log_likelihoods()[0], ll_a):
            assert np.array_equal(x, y)
        a2.refine(True)
        for x, y in zip(a2.events()[0], ev_a):
            assert np.array_equal(x, y)


def test_empty_and_degenerate_batches(lib_built):
    """Ragged corners: an empty batch, a read with an empty reference next to normal reads."""
    from nadavca_b200 import dtw
    rng = np.random.default_rng(31)
    k, cp, mel, bw = 2, 1, 2, 6
    mean = rng.normal(0, 1.2, size=16)
    sigma = rng.uniform(0.2, 0.6, size=16)
    gm = dtw.KmerModel(k, cp, 4, mean, sigma)
    with dtw.Batch(gm, [], [], [], [], [], bw, mel) as batch:
        batch.refine(True)
        assert batch.events()[0] == []
        batch.estimate(True)
        assert batch.log_likelihoods()[0] == []
    good = make_case(rng, k, cp, 25, bw, mel)
    lists = [[good[i], np.zeros(0)] for i in (2, 3, 4, 5)] + [[good[6], np.zeros((0, 2), dtype=int)]]
    lists[0][1] = rng.normal(0, 1, 7)  # a signal without any reference base
    with dtw.Batch(gm, *lists, bw, mel) as batch:
        batch.refine(False)
        events, status = batch.events()
        assert status.tolist() == [0, 2] and events[1] is None and events[0] is not None
        batch.estimate(True)
        lls, status = batch.log_likelihoods()
        assert status.tolist() == [0, 2] and lls[1].shape == (0, 4) and np.all(np.isfinite(lls[0]))


@pytest.mark.parametrize('mel', [4, 6])
def test_long_minimum_event_lengths(lib_built, mel):
    from nadavca_b200 import dtw
    from oracle import oracle as orc
    rng = np.random.default_rng(40 + mel)
    k, cp, bw = 3, 1, 12
    mean = rng.normal(0, 1.2, size=64)
    sigma = rng.uniform(0.2, 0.6, size=64)
    cases = [make_case(rng, k, cp, int(rng.integers(15, 50)), bw, mel, spacing=9) for _ in range(4)]
    _compare_batch(rng, [(mean, sigma) + c[2:] for c in cases], k, cp, mel, bw, mean, sigma)
    gm = dtw.KmerModel(k, cp, 4, mean, sigma)
    with pytest.raises(dtw.NadavcaCudaError):
        dtw.Batch(gm, [cases[0][2]], [cases[0][3]], [[]], [[]], [cases[0][6]], bw, 7)


@pytest.mark.parametrize('k,cp', [(8, 3), (10, 4)])
def test_large_kmer_models(lib_built, k, cp):
    """k-mer tables beyond the shipped 6-mer (the reference's default model is a 10-mer, 4^10 entries = 25 MB of
    tables): k+2 lanes per SNP task, 3 (k=8) or 2 (k=10) tasks per warp."""
    rng = np.random.default_rng(50 + k)
    mel, bw = 2, 10
    mean = rng.normal(0, 1.2, size=4 ** k)
    sigma = rng.uniform(0.25, 0.5, size=4 ** k)
    cases = [make_case(rng, 3, 1, int(rng.integers(20, 45)), bw, mel) for _ in range(3)]

    def with_model(c):
        # make_case built the signal from its own small model; re-simulate the levels from the large one
        ref = c[3]
        n = len(ref)
        padded = np.zeros(n + k, dtype=int)
        padded[cp:cp + n] = ref
        ids = np.zeros(n, dtype=int)
        for j in range(k):
            ids = ids * 4 + padded[j:j + n]
        lengths = np.maximum(mel, rng.poisson(6, size=n))
        starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]) + bw
        sig = np.concatenate([rng.normal(0, 1, bw), np.repeat(mean[ids], lengths) + rng.normal(0, 0.3, lengths.sum()),
                              rng.normal(0, 1, bw)])
        anchors = np.stack([np.clip(starts + rng.integers(-2, 3, size=n), 0, len(sig) - 1), np.arange(n)], axis=1)
        anchors[:, 0] = np.maximum.accumulate(anchors[:, 0])
        cb = rng.integers(0, 4, size=cp)
        ca = rng.integers(0, 4, size=k - cp - 1)
        return (mean, sigma, np.clip(sig, -5, 5), ref, cb, ca, anchors)
    _compare_batch(rng, [with_model(c) for c in cases], k, cp, mel, bw, mean, sigma)


@pytest.fixture
def sweep_schedule(request):
    """Force one of the two sweep schedules (the library picks by batch size otherwise): 'r' = rotating wavefront
    (rows5.cu), 's' = pipelined stripes (rows4.cu)."""
    from nadavca_b200 import dtw
    dtw.set_sweep_schedule(request.param)
    yield request.param
    dtw.set_sweep_schedule(None)


@pytest.mark.parametrize('sweep_schedule', ['r', 's'], indirect=True)
def test_both_sweep_schedules_match_oracle(lib_built, default_model, sweep_schedule):
    """Every mode through both schedules: random ragged batches (sparse anchors, several minimum event lengths, reads
    shorter and longer than one 32-pair generation) and full-size 6-mer reads on both strands."""
    for k, cp, mel in ((3, 1, 2), (4, 2, 1), (2, 0, 3), (6, 2, 2)):
        rng = np.random.default_rng(900 + 10 * k + mel)
        bw = int(rng.integers(4, 24))
        mean = rng.normal(0, 1.2, size=4 ** k)
        sigma = rng.uniform(0.2, 0.6, size=4 ** k)
        cases = []
        for i in range(8):
            n = int(rng.integers(1, 140)) if i else 1
            c = make_case(rng, k, cp, n, bw, mel, sparse=i % 3 == 1)
            cases.append((mean, sigma) + c[2:])
        _compare_batch(rng, cases, k, cp, mel, bw, mean, sigma)
    test_default_model_read_matches_oracle(default_model)


def test_very_wide_band_row_inside_a_batch(lib_built):
    """A basecaller stall: an anchor gap of 12 000 samples gives band rows wider than the shared-memory hand-off
    rows of the striped sweep (~8.5 k columns); the read goes through the global hand-off rows and the other reads
    of the batch are unaffected (the reference handles any band width)."""
    from nadavca_b200 import dtw
    from oracle import oracle as orc
    rng = np.random.default_rng(61)
    k, cp, mel, bw = 3, 1, 2, 20
    mean = rng.normal(0, 1.2, size=64)
    sigma = rng.uniform(0.25, 0.5, size=64)
    normal = [make_case(rng, k, cp, 40, bw, mel)[2:] for _ in range(2)]
    # the stalled read: 30 bases, the 13th event lasts 12 000 samples, anchors only outside the stall
    n = 30
    ref = rng.integers(0, 4, size=n)
    padded = np.zeros(n + k, dtype=int)
    padded[cp:cp + n] = ref
    ids = np.zeros(n, dtype=int)
    for j in range(k):
        ids = ids * 4 + padded[j:j + n]
    lengths = np.maximum(mel, rng.poisson(7, size=n))
    lengths[12] = 12_000
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]) + bw
    sig = np.clip(np.concatenate([rng.normal(0, 1, bw), np.repeat(mean[ids], lengths) +
                                  rng.normal(0, 0.35, lengths.sum()), rng.normal(0, 1, bw)]), -5, 5)
    keep = np.array([j for j in range(n) if j not in (11, 12, 13)])
    anchors = np.stack([starts[keep], keep], axis=1)
    stalled = (sig, ref, [], [], anchors)
    cases = [normal[0], stalled, normal[1]]
    gm = dtw.KmerModel(k, cp, 4, mean, sigma)
    om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
    with dtw.Batch(gm, *[[c[i] for c in cases] for i in range(5)], bw, mel) as batch:
        widest = max(int((be - bs + 1).max()) for bs, be in batch.bands())
        assert widest > 10_000
        for flag in (False, True):
            batch.refine(flag)
            events, status = batch.events()
            assert status.tolist() == [0, 0, 0]
            for ev, c in zip(events, cases):
                assert_same_path(ev, (mean, sigma) + tuple(c), bw, mel, om, flag)
        batch.estimate(True)
        lls, _ = batch.log_likelihoods()
        for ll, c in zip(lls, cases):
            want = np.array(orc.estimate_log_likelihoods(*c, bw, mel, om, True))
            np.testing.assert_allclose(ll, want, rtol=LL_RTOL, atol=LL_ATOL)


def test_batches_of_one_model_overlap_on_two_streams(lib_built):
    """Every stream has its own DP workspace: two batches of ONE model in flight on two streams give exactly the
    results of running them one after the other."""
    import torch
    from nadavca_b200 import dtw
    rng = np.random.default_rng(71)
    k, cp, mel, bw = 4, 2, 2, 14
    mean = rng.normal(0, 1.2, size=256)
    sigma = rng.uniform(0.2, 0.6, size=256)
    gm = dtw.KmerModel(k, cp, 4, mean, sigma)
    sets = [[make_case(rng, k, cp, int(rng.integers(80, 200)), bw, mel)[2:] for _ in range(24)] for _ in range(2)]
    lists = [[[c[i] for c in cases] for i in range(5)] for cases in sets]
    want = []
    for ls in lists:
        with dtw.Batch(gm, *ls, bw, mel) as b:
            b.refine(True)
            ev = [e.copy() for e in b.events()[0]]
            b.estimate(True)
            want.append((ev, [x.copy() for x in b.log_likelihoods()[0]]))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with dtw.Batch(gm, *lists[0], bw, mel) as a, dtw.Batch(gm, *lists[1], bw, mel) as b:
        for rep in range(3):
            a.refine(True, s1)
            b.refine(True, s2)
            a.estimate(True, s1)
            b.estimate(True, s2)
        for batch, (ev, ll) in ((a, want[0]), (b, want[1])):
            for x, y in zip(batch.events()[0], ev):
                assert np.array_equal(x, y)
            for x, y in zip(batch.log_likelihoods()[0], ll):
                assert np.array_equal(x, y)
