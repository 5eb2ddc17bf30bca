"""Parity bookkeeping shared by tests/, tools/ and bench.py's CPU leg.   *** TEST INFRASTRUCTURE ***

The bar for alignments is bit-exact event boundaries.  One structural exception exists and is DETECTED, not assumed:
when two consecutive reference positions carry the same emission (identical k-mer ids: a homopolymer run longer than
k, or any repeated base under a 1-mer model) every split of their samples has the same likelihood, the posterior
cells of the boundary between them tie mathematically, and the reference's own argmax (node.cpp:49-56,72-75,82-85,
strict '>') is decided by the last-ulp rounding of ITS log-space arithmetic.  Measured here with the oracle itself:
the plain-C port compiled with FMA contraction (different rounding, same algorithm) changes the path of 14 / 3268
random reads, every one of them at a boundary between two rows with identical k-mers and none anywhere else
(profiles/r02_tie_detector.txt).  So a path that differs from the oracle's is accepted only if
  (1) every differing row lies in `tie_rows` (rows adjacent to an identical-emission neighbour), and
  (2) it is feasible and its max-product score under the ORACLE's posterior rows equals the oracle path's score,
and the number of such accepts is counted and reported, never silent.
"""
import numpy as np

from . import oracle as orc


def kmer_ids(reference, context_before, context_after, k, central):
    """KmerModel::GetKmerId at every reference position with 'A' padding outside the contexts (kmer_model.cpp:22-30,
    sequence.cpp:23-28)."""
    ref = np.asarray(reference, dtype=np.int64).reshape(-1)
    cb = np.asarray(context_before, dtype=np.int64).reshape(-1)
    ca = np.asarray(context_after, dtype=np.int64).reshape(-1)
    n = len(ref)
    seq = np.concatenate([cb, ref, ca])
    padded = np.zeros(len(seq) + 2 * k, dtype=np.int64)  # k bases of 'A' (0) on both sides
    padded[k:k + len(seq)] = seq
    ids = np.zeros(n, dtype=np.int64)
    for t in range(k):
        lo = k + len(cb) - central + t  # reference position i reads padded[lo + i]
        ids = ids * 4 + padded[lo:lo + n]
    return ids


def tie_rows(reference, context_before, context_after, k, central, mean, sigma):
    """Boolean mask over the reference positions: True where the position has a neighbour with an identical emission
    (same mean and sigma), i.e. where the boundary towards that neighbour is a mathematical tie."""
    ids = kmer_ids(reference, context_before, context_after, k, central)
    mean = np.asarray(mean)
    sigma = np.asarray(sigma)
    same = (mean[ids[1:]] == mean[ids[:-1]]) & (sigma[ids[1:]] == sigma[ids[:-1]])
    mask = np.zeros(len(ids), dtype=bool)
    mask[1:] |= same
    mask[:-1] |= same
    return mask


def path_score(events, dbg, min_event_length, transitions):
    """Max-product score of a path (sum of the oracle's log posteriors prefix + suffix along it); asserts the path
    is inside the band and respects the minimum event lengths (dtw.cpp:165-179)."""
    bs, be = dbg['bs'], dbg['be']
    off = np.concatenate([[0], np.cumsum(be - bs + 1)])
    post = dbg['prefix'] + dbg['suffix']
    events = np.asarray(events)
    if transitions:
        cols = events.reshape(-1)
        mins = [min_event_length if r % 2 == 0 else 0 for r in range(len(cols) - 1)]
    else:
        cols = np.concatenate([events[:, 0], events[-1:, 1]])
        assert np.array_equal(events[1:, 0], events[:-1, 1]), 'events are not contiguous'
        mins = [min_event_length] * (len(cols) - 1)
    total = 0.0
    for r, c in enumerate(cols):
        assert bs[r] <= c <= be[r], 'path leaves the band at row %d' % r
        if r:
            assert c - cols[r - 1] >= mins[r - 1], 'event shorter than the minimum at row %d' % r
        total += post[off[r] + c - bs[r]]
    return total


def compare_events(ev, signal, reference, context_before, context_after, anchors, bandwidth, min_event_length, model,
                   transitions, want=None):
    """'exact' when the events equal the oracle's bit for bit, 'tie' when they differ only as described in the module
    docstring; raises AssertionError otherwise.  `ev` is an (n,2) int array or None (no path); `want` the oracle's
    events when already computed (e.g. by oracle/_ref in a worker process)."""
    if want is None:
        want = orc.refine_alignment(signal, reference, context_before, context_after, anchors, bandwidth,
                                    min_event_length, model, transitions)
    want = [list(map(int, w)) for w in want]
    if len(want) == 0:
        assert ev is None or len(ev) == 0, 'the oracle finds no path, the device did'
        return 'exact'
    assert ev is not None and len(ev) == len(want), 'the oracle finds a path, the device did not'
    got = np.asarray(ev).tolist()
    if got == want:
        return 'exact'
    mask = tie_rows(reference, context_before, context_after, model.k, model.central_position, model.mean,
                    model.sigma)
    rows = np.nonzero((np.asarray(got) != np.asarray(want)).any(axis=1))[0]
    assert mask[rows].all(), ('alignment differs from the oracle at rows %s, which have no identical-emission '
                              'neighbour (not a structural tie)' % rows[~mask[rows]][:8].tolist())
    port = model if model.backend == 'port' else orc.OracleModel(model.k, model.central_position,
                                                                 model.alphabet_size, model.mean, model.sigma, 'port')
    _, dbg = orc.refine_alignment(signal, reference, context_before, context_after, anchors, bandwidth,
                                  min_event_length, port, transitions, debug=True)
    a = path_score(got, dbg, min_event_length, transitions)
    b = path_score(want, dbg, min_event_length, transitions)
    assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), 'tie rows, but the path scores differ: %r vs %r' % (a, b)
    return 'tie'


def ll_max_rel(got, want):
    """max |got - want| / |want| over the finite entries; asserts that the finite patterns agree."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    finite = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), finite), 'finite patterns of the log-likelihoods differ'
    if not finite.any():
        return 0.0
    return float(np.max(np.abs(got[finite] - want[finite]) / np.maximum(np.abs(want[finite]), 1e-300)))


# ---- worker-process jobs (multiprocessing 'spawn': CUDA lives in the parent) ---------------------------------

_MODELS = {}


def _model(spec):
    k, cp, mean, sigma, backend = spec
    key = (k, cp, backend, mean.tobytes()[:64], len(mean))
    if key not in _MODELS:
        if backend == 'ref' and orc.ref_module() is None:
            backend = 'port'
        _MODELS[key] = orc.OracleModel(k, cp, 4, mean, sigma, backend)
    return _MODELS[key]


def job(task):
    """One oracle call for a pool worker.  task = (kind, model_spec, args) with args = (signal, reference, cb, ca,
    anchors, bandwidth, min_event_length, flag); kind 'refine' -> events list, 'estimate' -> (n,4) array."""
    kind, spec, args = task
    model = _model(spec)
    signal, reference, cb, ca, anchors, bw, mel, flag = args
    if kind == 'refine':
        return orc.refine_alignment(signal, reference, cb, ca, anchors, bw, mel, model, flag)
    if kind == 'estimate':
        return np.array(orc.estimate_log_likelihoods(signal, reference, cb, ca, anchors, bw, mel, model, flag))
    raise ValueError(kind)
