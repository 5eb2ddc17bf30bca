"""Both sweep schedules on the same batch: events must be identical, log-likelihoods equal to rounding.
  python tools/compare_schedules.py <reads> <bases> <bandwidth>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nadavca_b200 import dtw
reads, bases, bw = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
km = bench.load_model(); km._device = 0
torch.cuda.set_device(0)
genome, items = bench.make_workload(km, reads, 0, bases, max(1_000_000, 4 * bases), bw)
lists = ([it['signal'] for it in items], [it['reference'] for it in items], [it['cb'] for it in items],
         [it['ca'] for it in items], [it['apx'].alignment for it in items])
res = {}
with dtw.Batch(km, *lists, bw, 2) as b:
    for sched in ('s', 'r'):
        dtw.set_sweep_schedule(sched)
        out = {}
        for flag in (False, True):
            b.refine(flag); torch.cuda.synchronize(); t0 = time.perf_counter()
            b.refine(flag); ev, st = b.events(); out['refine%d_ms' % flag] = (time.perf_counter() - t0) * 1e3
            out['ev%d' % flag] = [e.copy() for e in ev]
        b.estimate(True); torch.cuda.synchronize(); t0 = time.perf_counter()
        b.estimate(True); ll, _ = b.log_likelihoods(); out['estimate_ms'] = (time.perf_counter() - t0) * 1e3
        out['ll'] = [x.copy() for x in ll]
        res[sched] = out
same_ev = all(np.array_equal(a, c) for f in (False, True) for a, c in zip(res['s']['ev%d' % f], res['r']['ev%d' % f]))
worst = max(float(np.max(np.abs(a - c) / np.abs(a))) for a, c in zip(res['s']['ll'], res['r']['ll']))
print('reads %d bases %d bw %d: events identical %s, max rel LL diff %.2e | stripes: refine %.0f / %.0f ms estimate %.0f ms | rotation: refine %.0f / %.0f ms estimate %.0f ms' %
      (reads, bases, bw, same_ev, worst, res['s']['refine0_ms'], res['s']['refine1_ms'], res['s']['estimate_ms'],
       res['r']['refine0_ms'], res['r']['refine1_ms'], res['r']['estimate_ms']))
