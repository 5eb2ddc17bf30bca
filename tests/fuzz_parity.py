"""Randomised parity sweep against the oracle (beyond the fixed seeds of tests/): python tests/fuzz_parity.py <seeds>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from conftest import make_case
import test_gpu_parity as T
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
ties = 0
t0 = time.time()
for seed in range(1000, 1000 + n_seeds):
    rng = np.random.default_rng(seed)
    k = int(rng.integers(2, 7)); cp = int(rng.integers(0, k)); mel = int(rng.integers(1, 5))
    bw = int(rng.integers(3, 60))
    mean = rng.normal(0, 1.2, size=4 ** k); sigma = rng.uniform(0.2, 0.6, size=4 ** k)
    cases = []
    for i in range(int(rng.integers(2, 10))):
        n = int(rng.integers(1, 220))
        c = make_case(rng, k, cp, n, bw, mel, sparse=bool(rng.integers(0, 2)), homopolymer=False,
                      spacing=int(rng.integers(max(mel, 2), 14)))
        cases.append((mean, sigma) + c[2:])
    before = T.TIE_LOG['tie']
    try:
        T._compare_batch(rng, cases, k, cp, mel, bw, mean, sigma)
        if T.TIE_LOG['tie'] > before:
            ties += T.TIE_LOG['tie'] - before
            print('seed', seed, 'k', k, 'cp', cp, 'mel', mel, 'bw', bw, 'structural tie(s):', T.TIE_LOG['tie'] - before)
    except AssertionError as e:
        bad += 1
        print('seed', seed, 'k', k, 'cp', cp, 'mel', mel, 'bw', bw, 'FAILED:', str(e)[:300].replace('\n', ' '))
print('%d seeds, %d failures, %d structural ties among %d alignments, %.0f s' % (n_seeds, bad, ties, sum(T.TIE_LOG.values()), time.time() - t0))
