import sys, os
os.environ['NVB_DEBUG_SKIP_PATH'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import make_case
from nadavca_b200 import dtw
from oracle import oracle as orc
k, cp, mel, n, bw = [int(x) for x in sys.argv[1:6]]
rng = np.random.default_rng(1)
mean, sigma, sig, ref, cb, ca, anc = make_case(rng, k, cp, n, bw, mel)
gm = dtw.KmerModel(k, cp, 4, mean, sigma)
om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
want, dbg = orc.refine_alignment(sig, ref, cb, ca, anc, bw, mel, om, True, debug=True)
bs, be = dbg['bs'], dbg['be']
off = np.concatenate([[0], np.cumsum(be - bs + 1)])
with dtw.Batch(gm, [sig], [ref], [cb], [ca], [anc], bw, mel) as batch:
    batch.refine(True)
    for plane, name in ((0, 'prefix'), (1, 'suffix')):
        got = batch.debug_rows(0, plane, transitions=True)
        exp = dbg[name]
        both = np.isfinite(got) & np.isfinite(exp)
        mism_inf = np.nonzero(np.isfinite(got) != np.isfinite(exp))[0]
        err = np.zeros_like(got); err[both] = np.abs(got[both] - exp[both])
        print(name, 'cells', len(got), 'inf-mismatch', len(mism_inf), 'max abs err', err.max())
        bad = np.union1d(np.nonzero(err > 1e-7)[0], mism_inf)
        if len(bad):
            rows = np.searchsorted(off, bad, side='right') - 1
            print('  bad rows', sorted(set(rows.tolist()))[:20])
            for b in bad[:6]:
                r = np.searchsorted(off, b, side='right') - 1
                print('   row', r, 'col', bs[r] + b - off[r], 'band', bs[r], be[r], 'got', got[b], 'want', exp[b])
