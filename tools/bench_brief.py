"""Run bench.py with the given arguments and print the step time and the per-stage milliseconds."""
import json, subprocess, sys
out = subprocess.run([sys.executable, 'bench.py', '--no-cpu-baseline'] + sys.argv[1:], capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith('{')]
if not line:
    print(out.stdout[-2000:], out.stderr[-2000:]); sys.exit(1)
d = json.loads(line[-1])
print('ms/step %.2f  value %.1fM  e2e %.1fM  clocks %s stages %s' % (d['ms_per_step'], d['value'] / 1e6, d['e2e']['value'] / 1e6, d['clocks'],
      {k: round(v, 2) for k, v in d['stage_ms_per_step'].items()}))
