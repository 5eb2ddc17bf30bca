"""Summarise bench JSON lines: python tools/bench_show.py gpurun_out/x.json [...]"""
import json, sys
for path in sys.argv[1:]:
    try:
        lines = [l for l in open(path).read().splitlines() if l.startswith('{')]
        d = json.loads(lines[-1])
    except Exception as e:
        print(path, 'unreadable:', e)
        continue
    out = ['%s: ms/step %.2f value %.1fM' % (path.split('/')[-1], d.get('ms_per_step', 0), d.get('value', 0) / 1e6)]
    if 'e2e' in d:
        out.append('e2e %.1fM' % (d['e2e']['value'] / 1e6))
    if d.get('stage_ms_per_step'):
        out.append('stages %s' % {k: round(v, 2) for k, v in d['stage_ms_per_step'].items()})
    if d.get('consensus'):
        c = d['consensus']
        out.append('consensus %.1f ms (%s, exchange %.3f ms, bus %s GB/s, identical %s, modes diff %.1e)' % (
            c['ms_per_step'], c['collective'], c['exchange_ms_per_step'], c.get('bus_GBs'), c['ranks_identical'],
            c['reduce_scatter_vs_allreduce_max_abs_diff']))
    if d.get('api'):
        a = d['api']
        if 'estimate_snps' in a:
            out.append('api estimate_snps %.3f s align_signal %.3f s' % (a['estimate_snps']['seconds'], a['align_signal']['seconds']))
        else:
            out.append('api %.3f s' % a['seconds'])
    if d.get('alu'):
        out.append('alu issue %.2f fp64 %.2f' % (d['alu'].get('issue_frac') or 0, d['alu'].get('fp64_pipe_frac') or 0))
    if d.get('parity'):
        out.append('parity %s' % d['parity'])
    if d.get('cpu_baseline'):
        out.append('cpu %.3fM (%s cores)' % (d['cpu_baseline']['value'] / 1e6, d['cpu_baseline']['cores']))
    print(' | '.join(out))
