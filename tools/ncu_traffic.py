"""Record the measured DRAM traffic of ONE advance_kernel launch (an `ncu --set full` capture of `bench.py --steps K ...`) in
profiles/step_kernel_traffic.json under the key bench.py looks up: '<precision>:<envs>:<steps per launch>'.
Usage: python tools/ncu_traffic.py <rep> <precision> <envs> <steps> [source label]"""
import csv
import json
import subprocess
import sys
from pathlib import Path

rep, precision, envs, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
label = sys.argv[5] if len(sys.argv) > 5 else rep
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]


def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(',', ''))
    u = units[i].lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)


path = Path(__file__).resolve().parents[1] / 'profiles' / 'step_kernel_traffic.json'
data = json.loads(path.read_text()) if path.is_file() else {}
data[f'{precision}:{envs}:{steps}'] = {
    'dram_read_bytes': get('dram__bytes_read.sum'), 'dram_write_bytes': get('dram__bytes_write.sum'),
    'duration_us_under_ncu': float(vals[hdr.index('gpu__time_duration.sum')].replace(',', '')),
    'source': label, 'note': 'one launch; writes still dirty in the 126 MB L2 at kernel end are not counted by dram__bytes_write'}
path.write_text(json.dumps(data, indent=1))
print(json.dumps(data[f'{precision}:{envs}:{steps}']))
