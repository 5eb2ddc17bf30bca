import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from nadavca_b200 import dtw, synthetic
from nadavca_b200.kmer_model import KmerModel
from nadavca_b200.read import Read
from oracle import oracle as orc
km = KmerModel.load_from_hdf5(os.path.join(ROOT, 'nadavca_b200', 'default', 'kmer_model.hdf5'))
om = orc.OracleModel(km.get_k(), km.get_central_position(), 4, km.mean, km.sigma, 'port')
genome = synthetic.make_genome(3000, seed=1)
reads = [synthetic.make_read(genome, km, i, n_bases=300, strand=s, substitution_rate=0.02) for i, s in enumerate('+-')]
Read.normalize_reads(reads)
aligner = synthetic.SyntheticAligner(genome)
args = []
for r in reads:
    apx = aligner.get_signal_alignment(r, 150)
    s0, s1 = apx.signal_range
    ref = orc.to_numerical(apx.reference_part)
    a, b = apx.read_sequence_range
    args.append((r.normalized_signal[s0:s1], ref, orc.to_numerical(r.sequence[a - 2:a]), orc.to_numerical(r.sequence[b:b + 3]), apx.alignment))
order = [bool(int(x)) for x in sys.argv[1]]
with dtw.Batch(km, *[list(x) for x in zip(*args)], 150, 2) as batch:
    for (bs, be) in batch.bands():
        print('maxw', (be - bs + 1).max(), 'n', len(bs) - 1)
    for flag in order:
        batch.refine(flag)
        events, st = batch.events()
        for ev, a in zip(events, args):
            want = orc.refine_alignment(*a, 150, 2, om, flag)
            print('flag', flag, 'status', st, 'match', ev is not None and ev.tolist() == want)
if os.environ.get('NVB_DEBUG_SKIP_PATH'):
    with dtw.Batch(km, *[list(x) for x in zip(*args)], 150, 2) as batch:
        batch.refine(True)
        for ri, a in enumerate(args):
            want, dbg = orc.refine_alignment(*a, 150, 2, om, True, debug=True)
            bs, be = dbg['bs'], dbg['be']
            off = np.concatenate([[0], np.cumsum(be - bs + 1)])
            if ri != len(args) - 1:
                continue  # debug_rows reads the last wave; base offsets differ per read -> only check the last read
        # read index 1 is at mat_base[1]; debug_rows handles it
        for ri in range(len(args)):
            want, dbg = orc.refine_alignment(*args[ri], 150, 2, om, True, debug=True)
            bs, be = dbg['bs'], dbg['be']
            off = np.concatenate([[0], np.cumsum(be - bs + 1)])
            for plane, name in ((0, 'prefix'), (1, 'suffix')):
                got = batch.debug_rows(ri, plane, transitions=True)
                exp = dbg[name]
                both = np.isfinite(got) & np.isfinite(exp)
                mism_inf = np.nonzero(np.isfinite(got) != np.isfinite(exp))[0]
                err = np.zeros_like(got); err[both] = np.abs(got[both] - exp[both])
                print('read', ri, name, 'cells', len(got), 'inf-mismatch', len(mism_inf), 'max abs err', err.max(), 'nan', int(np.isnan(got).sum()))
                if ri == 1 and plane == 0:
                    for r in range(100, 135):
                        seg = slice(off[r], off[r + 1])
                        g_, e_ = got[seg], exp[seg]
                        d = np.abs(np.where(np.isfinite(g_) & np.isfinite(e_), g_ - e_, 0))
                        badc = np.nonzero((d > 1e-7) | (np.isfinite(g_) != np.isfinite(e_)))[0]
                        print('ROW', r, 'band', bs[r], be[r], 'first bad col', (bs[r] + badc[0]) if len(badc) else None, 'n bad', len(badc))
                    for r in range(114, 114):
                        cols = range(max(bs[r], 876), min(be[r], 886) + 1)
                        print('row', r, 'band', bs[r], be[r], ' '.join('%d:%.6f/%.6f' % (c, got[off[r] + c - bs[r]], exp[off[r] + c - bs[r]]) for c in cols))
                bad = np.union1d(np.nonzero(err > 1e-6)[0], mism_inf)
                if len(bad):
                    rows = np.searchsorted(off, bad, side='right') - 1
                    print('  bad rows', sorted(set(rows.tolist()))[:20])
                    for b in bad[:6]:
                        r = np.searchsorted(off, b, side='right') - 1
                        print('   row', r, 'col', bs[r] + b - off[r], 'band', bs[r], be[r], 'got', got[b], 'want', exp[b])
