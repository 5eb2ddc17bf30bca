"""Throughput of the device hot path on the other BASELINE.json configurations (not the bench line): band-width
sweep (configs[4]), long reads (configs[3]) and the align_signal shape (configs[0]: refine with transitions).

  python tools/bench_configs.py <reads> <bases> <bandwidth> [steps]

Prints one line: samples/s and DP cells/s of refine(transitions), refine(plain) and estimate(wobbling), device-resident.
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nadavca_b200 import dtw

reads, bases, bw = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
km = bench.load_model(); km._device = 0
torch.cuda.set_device(0)
t0 = time.perf_counter()
genome, items = bench.make_workload(km, reads, 0, bases, max(1_000_000, 4 * bases), bw)
lists = ([it['signal'] for it in items], [it['reference'] for it in items], [it['cb'] for it in items],
         [it['ca'] for it in items], [it['apx'].alignment for it in items])
prep = time.perf_counter() - t0
with dtw.Batch(km, *lists, bw, 2) as b:
    samples = b.pack.total_signal
    cells = b.cell_counts(True)
    maxw = max(int((be - bs + 1).max()) for bs, be in b.bands())
    out = []
    for name, fn, key in (('refine(transitions)', lambda: b.refine(True), ['refine_transitions']),
                          ('refine(plain)', lambda: b.refine(False), ['refine_plain']),
                          ('estimate(wobbling)', lambda: b.estimate(True), ['estimate_fb', 'estimate_snp'])):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        out.append('%s %.1f ms %.1f Msamples/s %.1f Gcells/s' % (name, dt * 1e3, samples / dt / 1e6,
                                                                sum(cells[k] for k in key) / dt / 1e9))
    ev, st = b.events()
    ok = int((st == 0).sum())
print('reads %d bases %d bandwidth %d (max band row %d) samples %.1fM paths %d/%d | %s' %
      (reads, bases, bw, maxw, samples / 1e6, ok, reads, ' | '.join(out)))
