"""Host-side timing of the calls of one bench step (debugging aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nadavca_b200 import dtw
from nadavca_b200.estimator import ProbabilityEstimator
km = bench.load_model(); km._device = 0
torch.cuda.set_device(0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
genome, items = bench.make_workload(km, R, 0, 2000, 1_000_000, 150)
cfg = dict(bench.DEFAULT_CONFIG); est = ProbabilityEstimator(km, None, cfg)
lists = ([it['signal'] for it in items], [it['reference'] for it in items], [it['cb'] for it in items],
         [it['ca'] for it in items], [it['apx'].alignment for it in items])
reverse = [int(it['apx'].reverse_complement) for it in items]
intervals = [tuple(it['apx'].reference_range) for it in items]
stream = torch.cuda.current_stream()
bn = dtw.Batch(km, *lists, 150, 2); bt = dtw.Batch(km, *lists, 150, 2)
plan = est.plan_groups(intervals, genome, independent=True)
def sync(): torch.cuda.synchronize()
for it in range(4):
    t = [time.perf_counter()]
    bn.refine(False, stream); t.append(time.perf_counter()); sync(); t.append(time.perf_counter())
    bt.estimate(True, stream); t.append(time.perf_counter()); sync(); t.append(time.perf_counter())
    est.posterior_stage(bt, reverse, intervals, genome, independent=True, plan=plan); t.append(time.perf_counter()); sync(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print('refine call %.1f +sync %.1f | estimate call %.1f +sync %.1f | posterior call %.1f +sync %.1f' % tuple(d))
