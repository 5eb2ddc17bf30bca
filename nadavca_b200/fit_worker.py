"""Worker process of the spline-fit pool (``python -m nadavca_b200.fit_worker``): reads pickled lists of
(event means, expected levels) jobs from stdin, writes the pickled list of FITPACK splines to stdout.  A plain
subprocess, so it never imports the caller's ``__main__`` and never sees its CUDA context."""
import pickle
import struct
import sys


def main():
    from nadavca_b200.read import fit_spline
    inp, out = sys.stdin.buffer, sys.stdout.buffer
    while True:
        head = inp.read(8)
        if len(head) < 8:
            return
        (size,) = struct.unpack('<q', head)
        jobs = pickle.loads(inp.read(size))
        try:
            result = ('ok', [fit_spline(*job) for job in jobs])
        except Exception as exc:  # reported to the parent, which re-raises
            result = ('error', repr(exc))
        blob = pickle.dumps(result, protocol=pickle.HIGHEST_PROTOCOL)
        out.write(struct.pack('<q', len(blob)))
        out.write(blob)
        out.flush()


if __name__ == '__main__':
    main()
