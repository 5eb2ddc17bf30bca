import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import make_case
from nadavca_b200 import dtw
from oracle import oracle as orc
k, cp, mel, n, bw, nreads = [int(x) for x in sys.argv[1:7]]
rng = np.random.default_rng(1)
mean = rng.normal(0, 1.2, size=4 ** k); sigma = rng.uniform(0.2, 0.6, size=4 ** k)
cases = [make_case(rng, k, cp, n + 3 * i, bw, mel) for i in range(nreads)]
gm = dtw.KmerModel(k, cp, 4, mean, sigma)
om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
lists = [[c[i] for c in cases] for i in (2, 3, 4, 5, 6)]
with dtw.Batch(gm, *lists, bw, mel) as batch:
    for flag in (False, True):
        batch.refine(flag)
        ev, st = batch.events()
        print('flag', flag, 'status', st)
        for i, c in enumerate(cases):
            want = orc.refine_alignment(c[2], c[3], c[4], c[5], c[6], bw, mel, om, flag)
            got = None if ev[i] is None else ev[i].tolist()
            print('  read', i, 'match', got == want, (got[:3] if got else got), want[:3])
