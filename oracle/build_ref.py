"""Recipe for `oracle/_ref/`: the UNMODIFIED reference installed from `/root/reference` so that `bench.py --impl reference` and the
`cpu_baseline` leg can time the reference's own Python step on the bench host.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Nothing under `oracle/_ref/` is tracked (it is git-ignored) and the product never imports
it; it travels to the GPU box with the snapshot like a built `.so`.  Runs only where `/root/reference` exists (the build container):

    python oracle/build_ref.py

What it does: (1) `pip install --no-index --no-deps --no-build-isolation --target oracle/_ref/site <copy of /root/reference>` (the copy
lives under /tmp because the source tree is read-only); (2) copies the dataset directory the bench uses and the two files of
`data/misc` that `CityLearnEnv._load` asks `DataSet()` for (`citylearn/citylearn.py:2055-2057`) into `oracle/_ref/data`.  The two
pure-Python packages the reference imports but this image lacks (`gymnasium`, `simplejson`) are provided by the stand-ins in
`oracle/shims/`, exactly as for the golden fixtures.
"""
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path('/root/reference')
OUT = HERE / '_ref'
DATASETS = ['citylearn_challenge_2022_phase_all']


def build() -> bool:
    if not REF.is_dir():
        return False
    if OUT.exists():
        shutil.rmtree(OUT)
    (OUT / 'site').mkdir(parents=True)
    with tempfile.TemporaryDirectory(prefix='citylearn_ref_src_') as tmp:
        src = Path(tmp) / 'reference'
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns('docs', 'assets', 'examples', 'tests', '.git', 'data'))
        (src / 'data').mkdir()
        subprocess.run([sys.executable, '-m', 'pip', 'install', '--quiet', '--no-index', '--no-deps', '--no-build-isolation',
                        '--find-links', '/opt/wheelhouse', '--target', str(OUT / 'site'), str(src)], check=True)
    for ds in DATASETS:
        shutil.copytree(REF / 'data' / 'datasets' / ds, OUT / 'data' / 'datasets' / ds)
    (OUT / 'data' / 'misc').mkdir(parents=True)
    for f in ('battery_choices.yaml', 'lbl-tracking_the_sun-res-pv.csv'):
        shutil.copy(REF / 'data' / 'misc' / f, OUT / 'data' / 'misc' / f)
    return True


if __name__ == '__main__':
    print('oracle/_ref built' if build() else '/root/reference not present: nothing to do')
