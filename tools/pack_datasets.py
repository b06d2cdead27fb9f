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
    'baeda_3dem',
    'citylearn_challenge_2020_climate_zone_1', 'citylearn_challenge_2020_climate_zone_2', 'citylearn_challenge_2020_climate_zone_3',
    'citylearn_challenge_2020_climate_zone_4', 'citylearn_challenge_2021',
    'citylearn_challenge_2022_phase_1', 'citylearn_challenge_2022_phase_2', 'citylearn_challenge_2022_phase_3', 'citylearn_challenge_2022_phase_all',
    'citylearn_challenge_2023_phase_1', 'citylearn_challenge_2023_phase_2_local_evaluation',
    'citylearn_challenge_2023_phase_2_online_evaluation_1', 'citylearn_challenge_2023_phase_2_online_evaluation_2',
    'citylearn_challenge_2023_phase_2_online_evaluation_3',
    'citylearn_challenge_2023_phase_3_1', 'citylearn_challenge_2023_phase_3_2', 'citylearn_challenge_2023_phase_3_3',
    'citylearn_challenge_2022_phase_all_plus_evs', 'citylearn_charging_constraints_demo',
]

if __name__ == '__main__':
    root = Path(sys.argv[1]) if len(sys.argv) > 1 else Path('/root/reference/data/datasets')
    names = sys.argv[2:] or DEFAULT
    out = Path(__file__).resolve().parents[1] / 'citylearn_b200' / 'datasets'
    out.mkdir(exist_ok=True)
    for n in names:
        write_pack(DirectorySource(root / n), out / f'{n}.npz')
        print(n, (out / f'{n}.npz').stat().st_size)
