"""Experiment: refine of batch k+1 overlapped with estimate of batch k on two streams / two workspaces."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nadavca_b200 import dtw
km = bench.load_model(); km._device = 0
km2 = bench.load_model(); km2._device = 0
torch.cuda.set_device(0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
genome, items = bench.make_workload(km, R, 0, 2000, 1_000_000, 150)
lists = ([it['signal'] for it in items], [it['reference'] for it in items], [it['cb'] for it in items],
         [it['ca'] for it in items], [it['apx'].alignment for it in items])
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
bn = dtw.Batch(km, *lists, 150, 2); bt = dtw.Batch(km2, *lists, 150, 2)
def run(overlap, steps=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps):
        bn.refine(False, s1 if overlap else s1)
        bt.estimate(True, s2 if overlap else s1)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3
for ov in (False, True, False, True):
    print('overlap' if ov else 'serial ', '%.1f ms/step' % run(ov))
