"""Debug helper: run one random-batch configuration on the GPU and save the events next to the oracle's."""
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
    c = make_case(rng, k, cp, n, bw, mel, sparse=i % 3 == 1, homopolymer=i % 4 == 2)
    cases.append(c)
gm = dtw.KmerModel(k, cp, 4, mean, sigma)
om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
lists = [[c[i] for c in cases] for i in (2, 3, 4, 5, 6)]
out = {}
with dtw.Batch(gm, *lists, bw, mel) as batch:
    for flag in (False, True):
        batch.refine(flag)
        ev, st = batch.events()
        for ci, (e, c) in enumerate(zip(ev, cases)):
            want = orc.refine_alignment(c[2], c[3], c[4], c[5], c[6], bw, mel, om, flag)
            got = [] if e is None else e.tolist()
            if got != want:
                d = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
                print('MISMATCH case', ci, 'flag', flag, 'n', len(c[3]), 'rows differing', d[:10], len(d))
                for i in d[:5]:
                    print('   ', i, got[i], want[i])
