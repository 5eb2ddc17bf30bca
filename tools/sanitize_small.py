"""A small batch through every DP entry point (both band widths, min_event_length 0 and 2): a quick smoke run, and
the driver for `compute-sanitizer --tool memcheck python tools/sanitize_small.py [reads] [bases]` where a sanitizer is
available (it is closed on the round-2 GPU pool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from nadavca_b200 import dtw


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    bases = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    km = bench.load_model(); km._device = 0
    torch.cuda.set_device(0)
    for bw in (40, 150):
        genome, items = bench.make_workload(km, reads, 0, bases, 100_000, bw)
        lists = ([it['signal'] for it in items], [it['reference'] for it in items], [it['cb'] for it in items],
                 [it['ca'] for it in items], [it['apx'].alignment for it in items])
        stream = torch.cuda.current_stream()
        for mel in (0, 2):
            with dtw.Batch(km, *lists, bw, mel) as b:
                b.refine(False, stream); ev, st = b.events()
                b.refine(True, stream); b.events()
                b.estimate(True, stream); ll, _ = b.log_likelihoods()
                b.estimate(False, stream); b.log_likelihoods()
                torch.cuda.synchronize()
                print('bw', bw, 'mel', mel, 'reads', reads, 'status', list(st)[:4], 'll rows', sum(len(x) for x in ll))


if __name__ == '__main__':
    main()
