import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import make_case
from nadavca_b200 import dtw
from oracle import oracle as orc
k, cp, mel, n, bw, nreads = [int(x) for x in sys.argv[1:7]]
rng = np.random.default_rng(1)
mean = rng.normal(0, 1.2, size=4 ** k)
sigma = rng.uniform(0.2, 0.6, size=4 ** k)
cases = [make_case(rng, k, cp, n + 7 * i, bw, mel, sparse=(i % 2 == 1))[2:] for i in range(nreads)]
gm = dtw.KmerModel(k, cp, 4, mean, sigma)
om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
lists = [[c[i] for c in cases] for i in range(5)]
with dtw.Batch(gm, *lists, bw, mel) as batch:
    for flag in (False, True):
        batch.refine(flag)
        ev, st = batch.events()
        for i, c in enumerate(cases):
            want = orc.refine_alignment(*c, bw, mel, om, flag)
            bs, be = orc.band_bounds(c[4], len(c[0]), len(c[1]), bw)
            print('maxw', int((be - bs + 1).max()), 'flag', flag, 'read', i, 'status', st[i], 'match', ev[i] is not None and ev[i].tolist() == want)
