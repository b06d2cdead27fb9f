"""The device division forms of unit_physics.cuh (`dvd`, `dvr` + `Divisor`) against `__ddiv_rn`, bit for bit (tests/gpu/division_check.cu)."""
import ctypes
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

LIB = Path(__file__).resolve().parent / 'gpu' / 'libdivision_check.so'
MODES = {0: 'physical magnitudes', 1: 'edge of the fast range', 2: 'all exponents, negative divisors', 3: 'exact quotients',
         4: 'zero numerators', 5: 'near rounding ties', 6: 'curve-like decimals'}


@pytest.mark.parametrize('mode', sorted(MODES))
def test_division_forms_are_correctly_rounded(mode):
    assert LIB.is_file(), f'{LIB} not built: run __graft_entry__.build()'
    lib = ctypes.CDLL(str(LIB))
    lib.division_check.restype = ctypes.c_longlong
    lib.division_check.argtypes = [ctypes.c_uint64, ctypes.c_long, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    first = (ctypes.c_double * 4)()
    for seed in (1, 0xC17E1EA2):
        bad = lib.division_check(seed, 400_000_000, mode, first)
        assert bad == 0, f'{MODES[mode]}: {bad} mismatches, first x={first[0]!r} y={first[1]!r} ref={first[2]!r} got={first[3]!r}'
