"""Extract the DRAM traffic of the dominant kernel from an `ncu --set full` report into profiles/snp_traffic.json.

  python tools/ncu_traffic.py <report.ncu-rep> <reads_per_gpu>
"""
import csv, json, os, subprocess, sys
rep, reads = sys.argv[1], int(sys.argv[2])
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
best = None
for vals in rows[2:]:
    if 'snp3_kernel' not in vals[hdr.index('Kernel Name')]:
        continue
    def get(name):
        i = hdr.index(name)
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
        return float(vals[i].replace(',', '')) * scale
    best = {'kernel': 'snp3_kernel', 'reads_per_gpu': reads, 'dram_bytes_read': get('dram__bytes_read.sum'),
            'dram_bytes_write': get('dram__bytes_write.sum'), 'source': os.path.basename(rep)}
    best['dram_bytes_per_launch'] = best['dram_bytes_read'] + best['dram_bytes_write']
    plain = lambda name: float(vals[hdr.index(name)].replace(',', ''))
    best['warp_instructions_per_launch'] = plain('smsp__inst_executed.sum')
    best['launch_ms'] = plain('gpu__time_duration.sum')
    best['issue_active_pct'] = plain('smsp__issue_active.avg.pct_of_peak_sustained_active')
    best['fp64_pipe_pct'] = plain('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')
    best['alu_pipe_pct'] = plain('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active')
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if best is None:
    sys.exit('no snp3_kernel launch in ' + rep)
json.dump(best, open(os.path.join(root, 'profiles', 'snp_traffic.json'), 'w'), indent=1)
print(best)
