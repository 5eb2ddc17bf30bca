"""Debug helper: compare the stored prefix/suffix rows of the GPU with the oracle's for one random case."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import make_case
from nadavca_b200 import dtw
from oracle import oracle as orc

k, cp, mel, n, bw, wob = [int(x) for x in sys.argv[1:7]]
rng = np.random.default_rng(1)
mean, sigma, sig, ref, cb, ca, anc = make_case(rng, k, cp, n, bw, mel)
gm = dtw.KmerModel(k, cp, 4, mean, sigma)
om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
want, dbg = orc.estimate_log_likelihoods(sig, ref, cb, ca, anc, bw, mel, om, bool(wob), debug=True)
with dtw.Batch(gm, [sig], [ref], [cb], [ca], [anc], bw, mel) as batch:
    batch.estimate(bool(wob))
    ll, _ = batch.log_likelihoods()
    bs, be = dbg['bs'], dbg['be']
    off = np.concatenate([[0], np.cumsum(be - bs + 1)])
    for plane, name in ((0, 'prefix'), (1, 'suffix')):
        got = batch.debug_rows(0, plane)
        exp = dbg[name]
        both = np.isfinite(got) & np.isfinite(exp)
        mism_inf = np.nonzero(np.isfinite(got) != np.isfinite(exp))[0]
        err = np.zeros_like(got); err[both] = np.abs(got[both] - exp[both])
        print(name, 'cells', len(got), 'inf-mismatch', len(mism_inf), 'max abs err', err.max())
        bad = np.nonzero((err > 1e-7))[0]
        bad = np.union1d(bad, mism_inf)
        if len(bad):
            rows = np.searchsorted(off, bad, side='right') - 1
            print('  first bad rows', sorted(set(rows.tolist()))[:12])
            for b in bad[:6]:
                r = np.searchsorted(off, b, side='right') - 1
                print('   row', r, 'col', bs[r] + b - off[r], 'band', bs[r], be[r], 'got', got[b], 'want', exp[b])
    want = np.array(want)
    print('LL max abs err', np.nanmax(np.abs(ll[0] - want)))
    print(ll[0][:3], want[:3])
