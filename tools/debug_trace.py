import sys, os
os.environ['NVB_DEBUG_SKIP_PATH'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import make_case
from nadavca_b200 import dtw
k, cp, mel = 3, 1, 1
rng = np.random.default_rng(100 * k + mel)
bw = int(rng.integers(3, 20))
mean = rng.normal(0, 1.2, size=4 ** k); sigma = rng.uniform(0.2, 0.6, size=4 ** k)
cases = []
for i in range(12):
    n = int(rng.integers(1, 90)) if i else 1
    cases.append(make_case(rng, k, cp, n, bw, mel, sparse=i % 3 == 1, homopolymer=i % 4 == 2))
c = cases[3]
gm = dtw.KmerModel(k, cp, 4, mean, sigma)
with dtw.Batch(gm, [c[2]], [c[3]], [c[4]], [c[5]], [c[6]], bw, mel) as batch:
    batch.refine(True)
    batch.events
    import ctypes
    batch.debug_rows(0, 0, transitions=True)
