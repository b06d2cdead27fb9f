"""Pack reference datasets (schema.json + CSV + .pth) into the bundled `.npz` form.

Usage (in the build container, where /root/reference exists):
    python tools/pack_datasets.py [/root/reference/data/datasets] [name ...]
Data files only (public CityLearn datasets) - no reference source code is copied.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from citylearn_b200.data import DirectorySource, write_pack  # noqa: E402

DEFAULT = [
    'citylearn_challenge_2022_phase_1',
    'citylearn_challenge_2022_phase_all',
    'citylearn_challenge_2023_phase_2_local_evaluation',
]

if __name__ == '__main__':
    root = Path(sys.argv[1]) if len(sys.argv) > 1 else Path('/root/reference/data/datasets')
    names = sys.argv[2:] or DEFAULT
    out = Path(__file__).resolve().parents[1] / 'citylearn_b200' / 'datasets'
    out.mkdir(exist_ok=True)
    for n in names:
        write_pack(DirectorySource(root / n), out / f'{n}.npz')
        print(n, (out / f'{n}.npz').stat().st_size)
