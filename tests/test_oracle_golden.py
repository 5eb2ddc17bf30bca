"""CPU suite, part 1: the oracle is pinned against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py; the reference ships no tests or fixtures of its own).

  * the C restatement (oracle/nadavca_oracle.c, back end 'port') must reproduce the reference's
    refine_alignment / estimate_log_likelihoods / get_expected_signal outputs BIT FOR BIT;
  * when oracle/_ref (the compiled reference itself) is present it must reproduce them too -- this guards the
    fixtures against a stale or differently built _ref;
  * the Python restatement of the estimator glue (oracle.OracleEstimator) must reproduce the outputs of the
    reference's own nadavca/estimator.py bit for bit.
"""
import numpy as np
import pytest

from conftest import GOLDEN_CONFIG, golden_reads
from oracle import oracle as orc

BACKENDS = ['port'] + (['ref'] if orc.ref_module() is not None else [])


def _same_floats(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize('backend', BACKENDS)
def test_dp_cases_bit_exact(golden_dp, backend):
    names = [str(x) for x in golden_dp['names']]
    assert len(names) >= 50
    n_nopath = 0
    for tag in names:
        c = golden_dp.case(tag)
        k, cp, bw, mel = (int(x) for x in c['params'])
        model = orc.OracleModel(k, cp, 4, c['mean'], c['sigma'], backend)
        args = (c['signal'], c['reference'], c['context_before'], c['context_after'], c['anchors'], bw, mel, model)
        for flag in (0, 1):
            ev = orc.refine_alignment(*args, bool(flag))
            want = c['events%d' % flag]
            assert np.array_equal(np.asarray(ev, dtype=np.int32).reshape(-1, 2), want), (tag, flag)
            n_nopath += len(want) == 0
            ll = orc.estimate_log_likelihoods(*args, bool(flag))
            assert _same_floats(ll, c['ll%d' % flag]), (tag, flag)
        assert _same_floats(model.get_expected_signal(c['reference'], c['context_before'], c['context_after']),
                            c['expected']), tag
    assert n_nopath >= 2  # the fixtures contain infeasible bands (reference returns [])


def test_survey_known_answers(golden_dp):
    """The hand-checked vectors of SURVEY.md section 4 are what the golden file holds for the toy cases."""
    c = golden_dp.case('toy0')
    assert c['events0'].tolist() == c['events1'].tolist() == [[0, 3], [3, 6], [6, 9], [9, 11]]
    np.testing.assert_allclose(c['ll1'][0], [-0.013043763389, -4.261977157088, -16.055214132595, -35.692998888034],
                               rtol=0, atol=1e-11)
    np.testing.assert_allclose(c['ll1'][3], [-20.042425824037, -7.636074636067, -2.296276451577, -0.013043763389],
                               rtol=0, atol=1e-11)
    np.testing.assert_allclose(c['ll0'][0], [-0.500222199981, -1.982921519971, -7.528306532021, -19.725561807744],
                               rtol=0, atol=1e-11)
    assert golden_dp.case('toy1')['events1'].tolist() == [[0, 2], [2, 4], [4, 6], [6, 8]]
    assert golden_dp.case('toy2')['events1'].shape == (0, 2)


def test_band_bounds_properties(golden_dp):
    """Bands are inclusive, monotone and clipped to [0, N] (dtw.cpp:7-35)."""
    for tag in [str(x) for x in golden_dp['names']]:
        c = golden_dp.case(tag)
        k, cp, bw, mel = (int(x) for x in c['params'])
        n, N = len(c['reference']), len(c['signal'])
        bs, be = orc.band_bounds(c['anchors'], N, n, bw)
        assert len(bs) == len(be) == n + 1
        assert np.all(np.diff(bs) >= 0) and np.all(np.diff(be) >= 0)
        assert bs.min() >= 0 and be.max() <= N and be[-1] == N
        # rows with an anchor are centred on it
        for s, r in c['anchors']:
            assert bs[r] >= max(0, s - bw) and be[r] <= min(N, s + bw)


def test_events_properties(golden_dp):
    """Events are ordered, at least min_event_length long and contiguous without transitions."""
    for tag in [str(x) for x in golden_dp['names']]:
        c = golden_dp.case(tag)
        mel = int(c['params'][3])
        for flag in (0, 1):
            ev = c['events%d' % flag]
            if len(ev) == 0:
                continue
            assert np.all(ev[:, 1] - ev[:, 0] >= mel)
            assert np.all(ev[1:, 0] >= ev[:-1, 1])
            if not flag:
                assert np.array_equal(ev[1:, 0], ev[:-1, 1])
        # the reference-base column holds the no-SNP total: one value for the whole read (dtw.cpp:98-99)
        for flag in (0, 1):
            ll = c['ll%d' % flag]
            ref_col = ll[np.arange(len(ll)), c['reference']]
            assert np.all(ref_col == ref_col[0])


@pytest.mark.parametrize('tweak', [1, 0])
def test_estimator_glue_bit_exact(golden_estimator, default_model_host, tweak):
    """oracle.OracleEstimator == the reference's estimator.py on the stored reads."""
    from nadavca_b200 import synthetic
    g = golden_estimator
    km = default_model_host
    genome = g['genome']
    cfg = dict(GOLDEN_CONFIG, tweak_signal_normalization=bool(tweak))
    reads = golden_reads(g)
    from nadavca_b200.read import Read
    Read.normalize_reads(reads)
    for i, r in enumerate(reads):
        assert np.array_equal(r.normalized_signal, g['read%d/normalized_signal' % i])
    aligner = synthetic.SyntheticAligner(genome)
    om = orc.OracleModel(km.get_k(), km.get_central_position(), 4, km.mean, km.sigma, 'port')
    est = orc.OracleEstimator(om, aligner, cfg)
    pre = 'tweak%d/' % tweak
    for i, r in enumerate(reads):
        apx, table = est.get_refined_alignment(r)
        assert np.array_equal(table, g[pre + 'read%d/alignment_table' % i])
        chunk = est.estimate_log_likelihoods(genome, r)
        assert [chunk.start, chunk.end] == g[pre + 'read%d/chunk_range' % i].tolist()
        assert _same_floats(chunk.values, g[pre + 'read%d/chunk_values' % i])
        ind = est.estimate_probabilities(genome, [r])[0]
        assert _same_floats(ind.values, g[pre + 'read%d/independent_probabilities' % i])
    groups = est.estimate_probabilities(genome, reads)
    assert len(groups) == int(g[pre + 'n_groups']) == 4
    for gi, chunk in enumerate(groups):
        assert [chunk.start, chunk.end] == g[pre + 'group%d/range' % gi].tolist()
        assert _same_floats(chunk.values, g[pre + 'group%d/probabilities' % gi])
        assert np.array_equal(chunk.coverage, g[pre + 'group%d/coverage' % gi])
        np.testing.assert_allclose(chunk.values.sum(axis=1), 1.0, rtol=1e-12)


def test_cell_count_formulas(golden_dp):
    """count_cells against a direct restatement of the formulas from the band widths."""
    for tag in ['rnd05', 'rnd17', 'model6_0']:
        c = golden_dp.case(tag)
        k, cp, bw, mel = (int(x) for x in c['params'])
        n, N = len(c['reference']), len(c['signal'])
        bs, be = orc.band_bounds(c['anchors'], N, n, bw)
        W = (be - bs + 1).astype(int)
        got = orc.count_cells(c['anchors'], N, n, bw, k, cp)
        wt = [W[r // 2] if r % 2 == 0 else W[r // 2 + 1] for r in range(2 * n)]
        assert got['refine_transitions'] == sum(wt[1:]) + sum(wt[:2 * n - 1])
        assert got['refine_plain'] == W[1:].sum() + W[:-1].sum()
        assert got['estimate_fb'] == W[1:].sum() + W[:-1].sum() + 2 * W[1:n].sum()
        snp = 0
        for i in range(n):
            first, last = max(0, i - (k - cp - 1)), min(n - 1, i + cp)
            cells = sum(W[j + 1] + (W[j] if j > 0 else 0) for j in range(first, last + 1))
            cells += W[last] if last + 1 < n else 0
            snp += 3 * cells
        assert got['estimate_snp'] == snp
