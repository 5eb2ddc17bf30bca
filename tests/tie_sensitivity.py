"""How often does rounding noise alone change the reference's alignment, and where?  (CPU only.)

The plain-C oracle port is compiled twice: as pinned (no FMA contraction, bit-identical to the reference) and with
-O3 -march=native -ffp-contract=fast (same algorithm, different last-ulp rounding).  Every random read whose two
alignments differ is checked against oracle.parity.tie_rows: the differing rows must all sit next to a neighbour
with an identical emission.  This is the evidence behind the tie policy of the parity tests.

  python tests/tie_sensitivity.py [seeds] > profiles/r02_tie_detector.txt
"""
import os, pickle, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np

CHILD = r'''
import sys, os, pickle
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import numpy as np
from oracle import oracle as orc
if sys.argv[1] != 'pinned':
    orc.PORT_LIB = sys.argv[1]
from conftest import make_case
out = {}
for seed in range(1000, 1000 + int(sys.argv[2])):
    rng = np.random.default_rng(seed)
    k = int(rng.integers(1, 7)); cp = int(rng.integers(0, k)); mel = int(rng.integers(0, 5))
    bw = int(rng.integers(3, 60))
    mean = rng.normal(0, 1.2, size=4 ** k); sigma = rng.uniform(0.2, 0.6, size=4 ** k)
    om = orc.OracleModel(k, cp, 4, mean, sigma, 'port')
    for i in range(int(rng.integers(2, 10))):
        n = int(rng.integers(1, 220))
        c = make_case(rng, k, cp, n, bw, mel, sparse=bool(rng.integers(0, 2)), homopolymer=bool(rng.integers(0, 4) == 0),
                      spacing=int(rng.integers(max(mel, 2), 14)))
        for flag in (False, True):
            ev = orc.refine_alignment(c[2], c[3], c[4], c[5], c[6], bw, mel, om, flag)
            out[(seed, i, flag)] = (k, cp, mel, mean, sigma, c[3], c[4], c[5], ev)
pickle.dump(out, open(sys.argv[3], 'wb'))
'''

def main():
    seeds = sys.argv[1] if len(sys.argv) > 1 else '300'
    from oracle import parity
    with tempfile.TemporaryDirectory() as tmp:
        lib = os.path.join(tmp, 'libcontracted.so')
        subprocess.run(['gcc', '-O3', '-march=native', '-ffp-contract=fast', '-fPIC', '-shared', '-std=gnu11', '-o', lib,
                        os.path.join(ROOT, 'oracle', 'nadavca_oracle.c'), '-lm'], check=True)
        child = os.path.join(tmp, 'child.py')
        open(child, 'w').write(CHILD % {'root': ROOT})
        procs = [subprocess.Popen([sys.executable, child, which, seeds, os.path.join(tmp, name)])
                 for which, name in (('pinned', 'a.pkl'), (lib, 'b.pkl'))]
        for p in procs:
            assert p.wait() == 0
        a = pickle.load(open(os.path.join(tmp, 'a.pkl'), 'rb'))
        b = pickle.load(open(os.path.join(tmp, 'b.pkl'), 'rb'))
    total = flagged = differ = outside = 0
    for key, (k, cp, mel, mean, sigma, ref, cb, ca, ev) in a.items():
        mask = parity.tie_rows(ref, cb, ca, k, cp, mean, sigma)
        total += 1
        flagged += bool(mask.any())
        other = b[key][8]
        if ev != other:
            differ += 1
            if len(ev) != len(other):
                outside += 1
                print('read', key, 'k', k, 'mel', mel, 'one build finds a path, the other does not')
                continue
            rows = np.nonzero((np.array(ev) != np.array(other)).any(axis=1))[0]
            ok = bool(mask[rows].all())
            outside += not ok
            print('read', key, 'k', k, 'cp', cp, 'mel', mel, 'transitions', key[2], 'rows that differ', rows.tolist(),
                  'all tie rows' if ok else 'NOT ALL TIE ROWS')
    print('%d alignments (random k 1..6, min_event_length 0..4, both modes): %d reads contain tie rows, %d alignments '
          'change under FMA contraction, %d of those outside tie rows' % (total, flagged, differ, outside))

if __name__ == '__main__':
    main()
