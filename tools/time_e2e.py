"""Host-side timing of the pieces of bench.py's end-to-end step (debugging aid)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nadavca_b200 import _cabi, dtw
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
pk = _cabi.ReadsPack(*lists, 150, 2)
def pinned(arr):
    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory(); return t, t.numpy()
keep, host = [], {}
for name in ('signal', 'signal_off', 'reference', 'reference_off', 'context_before', 'context_before_off',
             'context_after', 'context_after_off', 'anchors', 'anchor_off'):
    t, a = pinned(getattr(pk, name)); keep.append(t); host[name] = a
from scipy import interpolate
xs = np.linspace(-4, 4, 40)
splines = [interpolate.splrep(xs, xs + 0.01 * np.sin(xs), s=40) for _ in range(pk.n_reads)]
pack = _cabi.ReadsPack.from_packed(bandwidth=150, min_event_length=2, **host)
ev_host = torch.empty((pk.total_reference, 2), dtype=torch.int32).pin_memory()
st_host = torch.empty(pk.n_reads, dtype=torch.int32).pin_memory()
prob_host = torch.empty((pk.total_reference, 4), dtype=torch.float64).pin_memory()
plan = est.plan_groups(intervals, genome, independent=True)
lib = _cabi.load()
for it in range(4):
    t = [time.perf_counter()]
    b = dtw.Batch.from_pack(km, pack); t.append(time.perf_counter())
    b.refine(False, stream); t.append(time.perf_counter())
    _cabi.check(lib.nvb_batch_get_events(b.handle, ctypes.cast(ev_host.data_ptr(), _cabi.c_i32p), ctypes.cast(st_host.data_ptr(), _cabi.c_i32p)), 'ev'); t.append(time.perf_counter())
    b.event_means(); t.append(time.perf_counter())
    b.apply_splines(splines, stream); t.append(time.perf_counter())
    b.estimate(True, stream); t.append(time.perf_counter())
    res = est.posterior_stage(b, reverse, intervals, genome, independent=True, plan=plan); t.append(time.perf_counter())
    prob_host.copy_(res[2], non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    b.close(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print('create %.1f | refine %.1f | get_events %.1f | event_means %.1f | apply_splines %.1f | estimate %.1f | posterior %.1f | d2h+sync %.1f | close %.1f | total %.1f' % (tuple(d) + (d.sum(),)))
