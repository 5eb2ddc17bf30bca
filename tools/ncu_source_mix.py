"""Per-source-line instruction and stall shares of one kernel out of an `ncu --set full --import-source on` report
(the view that found the round-2 savings: re-materialised constants, convergence tests, control flow around per-lane
conditions, dependent global loads).

  python tools/ncu_source_mix.py <report.ncu-rep> <kernel base-name regex> [top N] [substring of the full name]
      > profiles/<name>_source_mix.txt
(ncu matches --kernel-name against the base name; the substring picks one template instance, e.g. "(int)2, (int)0").
An instruction inlined from a header is listed under every file of its inline stack, so the totals count it more than
once: read the shares as relative weights, the kernel-level instruction counts are in the *_full.txt summaries.
"""
import csv, subprocess, sys


def main():
    rep, pattern = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    want = sys.argv[4] if len(sys.argv) > 4 else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                          'regex:' + pattern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines, ops, cur, hdr, kernel, take = {}, {}, None, None, None, False
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            take = want in r[1]
            if take and kernel is None:
                kernel = r[1]
            continue
        if r[0] == 'Line No':
            hdr = r
            ia, ist = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
            continue
        if hdr is None or len(r) <= ia or not take:
            continue
        try:
            n, s = int(r[ia]), int(r[ist])
        except ValueError:
            continue
        if r[0].isdigit():      # a CUDA source line: totals of the SASS lines below it
            key = (cur, int(r[0]))
            a = lines.setdefault(key, [0, 0, r[1].strip()])
            a[0] += n; a[1] += s
        elif r[2].startswith('0x'):  # one SASS instruction
            text = r[3].strip()
            if text.startswith('@'):
                text = text.split(None, 1)[1]
            op = text.split()[0].split('.')[0]
            b = ops.setdefault(op, [0, 0])
            b[0] += n; b[1] += s
    tot = sum(a[0] for a in lines.values()) or 1
    tots = sum(a[1] for a in lines.values()) or 1
    print('# %s' % kernel)
    print('# %s: warp instructions executed (all launches in the report) %.3f G, stall samples %d' % (rep.split('/')[-1], tot / 1e9, tots))
    print('\n## by source line (instruction share, stall-sample share)')
    for (f, l), (n, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top_n]:
        print('%-22s %4d  %5.2f%%  %5.2f%%  %s' % (f, l, 100 * n / tot, 100 * s / tots, src[:96]))
    otot = sum(b[0] for b in ops.values()) or 1
    print('\n## by opcode (instruction share, stall-sample share)')
    for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:24]:
        print('%-10s %5.2f%%  %5.2f%%' % (op, 100 * n / otot, 100 * s / tots))


if __name__ == '__main__':
    main()
