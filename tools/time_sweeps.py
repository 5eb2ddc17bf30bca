"""Stage times of refine(plain) / refine(transitions) / estimate(wobbling) on the bench workload, nothing else
(for A/B builds, including the NVB_EXPERIMENT_* timing experiments whose results are meaningless):
  python tools/time_sweeps.py [reads] [steps] [bandwidth]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from nadavca_b200 import dtw


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    bw = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    km = bench.load_model(); km._device = 0
    torch.cuda.set_device(0)
    genome, items = bench.make_workload(km, reads, 0, 2000, 4_600_000, bw)
    lists = ([it['signal'] for it in items], [it['reference'] for it in items], [it['cb'] for it in items],
             [it['ca'] for it in items], [it['apx'].alignment for it in items])
    stream = torch.cuda.current_stream()
    with dtw.Batch(km, *lists, bw, 2) as b:
        out = []
        for name, fn in (('refine(plain)', lambda: b.refine(False, stream)), ('refine(transitions)', lambda: b.refine(True, stream)),
                         ('estimate(wobbling)', lambda: b.estimate(True, stream))):
            fn(); torch.cuda.synchronize()
            b.enable_timing(True)
            for _ in range(steps):
                fn()
            t = b.timing()
            b.enable_timing(False)
            out.append('%s: %s' % (name, {k: round(v[0] / steps, 2) for k, v in t.items() if v[1]}))
        print('reads %d bw %d | %s' % (reads, bw, ' | '.join(out)))


if __name__ == '__main__':
    main()
