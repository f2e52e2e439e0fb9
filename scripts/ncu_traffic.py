"""DRAM traffic of the dominant kernels from committed `ncu --set full` captures -> profiles/r2_traffic.json
(bench.py reads it for roofline.traffic).  One entry per capture: kernel key, report, work units in the profiled launch.
   python scripts/ncu_traffic.py key=report.ncu-rep:units [...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
res = {}
if os.path.exists(out_path):
    with open(out_path) as f:
        res = json.load(f)
for arg in sys.argv[1:]:
    key, rest = arg.split('=', 1)
    rep, units = rest.rsplit(':', 1)
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit_row, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, unit_row))

    def to_bytes(name):
        v = float(d[name].replace(',', ''))
        return v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u[name]]
    rd, wr = to_bytes('dram__bytes_read.sum'), to_bytes('dram__bytes_write.sum')
    res[key] = dict(report=os.path.basename(rep), kernel=d.get('Kernel Name'), units_in_launch=float(units), dram_bytes_read=rd, dram_bytes_write=wr,
                    dram_bytes_per_unit=(rd + wr) / float(units), duration_ms=float(d['gpu__time_duration.sum'].replace(',', '')) * {'usecond': 1e-3, 'us': 1e-3, 'msecond': 1.0, 'ms': 1.0, 'second': 1e3, 's': 1e3, 'nsecond': 1e-6, 'ns': 1e-6}[u['gpu__time_duration.sum']])
    print(key, res[key])
with open(out_path, 'w') as f:
    json.dump(res, f, indent=1, sort_keys=True)
