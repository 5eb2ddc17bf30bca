import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import make_case
from nadavca_b200 import dtw
from oracle import oracle as orc
k, cp, mel, bw = 3, 1, 1, 6
for n in (30, 31, 32, 33, 61, 62, 63, 64, 80, 93, 94, 95, 130):
    rng = np.random.default_rng(1)
    mean, sigma, sig, ref, cb, ca, anc = make_case(rng, k, cp, n, bw, mel)
    gm = dtw.KmerModel(k, cp, 4, mean, sigma)
    om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
    res = []
    for flag in (True, False):
        want = orc.refine_alignment(sig, ref, cb, ca, anc, bw, mel, om, flag)
        with dtw.Batch(gm, [sig], [ref], [cb], [ca], [anc], bw, mel) as batch:
            batch.refine(flag)
            ev, st = batch.events()
        res.append((int(st[0]), None if ev[0] is None else ev[0].tolist() == want))
    print(n, res)
