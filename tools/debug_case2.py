import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import make_case
from nadavca_b200 import dtw
from oracle import oracle as orc
k, cp, mel = [int(x) for x in sys.argv[1:4]]
rng = np.random.default_rng(100 * k + mel)
bw = int(rng.integers(3, 20))
mean = rng.normal(0, 1.2, size=4 ** k); sigma = rng.uniform(0.2, 0.6, size=4 ** k)
cases = []
for i in range(12):
    n = int(rng.integers(1, 90)) if i else 1
    cases.append(make_case(rng, k, cp, n, bw, mel, sparse=i % 3 == 1, homopolymer=i % 4 == 2))
gm = dtw.KmerModel(k, cp, 4, mean, sigma)
om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
print('bw', bw, 'n', [len(c[3]) for c in cases], 'anchors', [len(c[6]) for c in cases])
def run(idx, flag):
    lists = [[cases[j][i] for j in idx] for i in (2, 3, 4, 5, 6)]
    with dtw.Batch(gm, *lists, bw, mel) as batch:
        batch.refine(flag)
        ev, st = batch.events()
    return [int(s) for s in st]
print('batch TRANS ', run(range(12), True))
print('single TRANS', [run([j], True)[0] for j in range(12)])
print('pairs  TRANS', [run([j, j], True) for j in (3, 6)])
print('prefix batches', [run(range(j + 1), True)[-1] for j in range(12)])
os.environ['NVB_DEBUG_SKIP_PATH'] = '1'
for ci in (3,):
    c = cases[ci]
    want, dbg = orc.refine_alignment(c[2], c[3], c[4], c[5], c[6], bw, mel, om, True, debug=True)
    bs, be = dbg['bs'], dbg['be']
    off = np.concatenate([[0], np.cumsum(be - bs + 1)])
    with dtw.Batch(gm, [c[2]], [c[3]], [c[4]], [c[5]], [c[6]], bw, mel) as batch:
        batch.refine(True)
        gp = batch.debug_rows(0, 0, transitions=True); gs = batch.debug_rows(0, 1, transitions=True)
    for name, got, exp in (('prefix', gp, dbg['prefix']), ('suffix', gs, dbg['suffix'])):
        both = np.isfinite(got) & np.isfinite(exp)
        mism = np.nonzero(np.isfinite(got) != np.isfinite(exp))[0]
        err = np.zeros_like(got); err[both] = np.abs(got[both] - exp[both])
        print(name, 'inf-mismatch', len(mism), 'maxerr', err.max())
        bad = np.union1d(np.nonzero(err > 1e-7)[0], mism)
        for b in bad[:8]:
            r = np.searchsorted(off, b, side='right') - 1
            print('   row', r, 'col', bs[r] + b - off[r], 'band', bs[r], be[r], 'got', got[b], 'want', exp[b])
    post_g = gp + gs; post_o = dbg['prefix'] + dbg['suffix']
    for r in range(len(bs)):
        a = post_g[off[r]:off[r+1]]; b = post_o[off[r]:off[r+1]]
        if np.isfinite(a).sum() != np.isfinite(b).sum():
            print('row', r, 'finite cells got', np.isfinite(a).sum(), 'want', np.isfinite(b).sum())
    print('ref around 60:', c[3][55:66])
    for r in (119, 120, 121):
        print('row', r, 'band', bs[r], be[r])
        print('  got ', np.round(gp[off[r]:off[r+1]], 3))
        print('  want', np.round(dbg['prefix'][off[r]:off[r+1]], 3))
    print('signal', np.round(c[2][355:375], 3))
    ids = [int(c[3][max(0,i-1)])*16 + int(c[3][i])*4 + int(c[3][min(len(c[3])-1,i+1)]) for i in (59, 60, 61)]
    print('kmer means', mean[ids], 'sigma', sigma[ids])
