"""The C header, the Python constants and the built library must agree (no GPU needed: nothing is launched)."""
import ctypes
import re
from pathlib import Path

import pytest

from citylearn_b200 import schema as S

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / 'include' / 'citylearn_b200.h').read_text()


def enum_members(name):
    body = re.search(r'enum\s+' + name + r'\s*\{(.*?)\}', HEADER, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    out, value = {}, 0
    for item in body.split(','):
        item = item.strip()
        if not item:
            continue
        if '=' in item:
            k, v = [x.strip() for x in item.split('=')]
            value = int(v, 0) if re.fullmatch(r'-?\w+', v) and not v.startswith('CL_') else int(eval(v, {}, dict(out)))
        else:
            k = item
        out[k] = value
        value += 1
    return out


def test_param_enums_match_python():
    p = enum_members('cl_building_param')
    assert p['CL_NPARAM'] == S.NPARAM
    for k, v in S.P.items():
        assert p['CL_P_' + k] == v, k
    ip = enum_members('cl_building_iparam')
    assert ip['CL_NIPARAM'] == S.NIPARAM
    for k, v in S.IP.items():
        assert ip['CL_IP_' + k] == v, k
    dyn = enum_members('cl_dyn')
    assert dyn['CL_NDYN'] == S.NDYN
    for k, v in S.DYN.items():
        assert dyn['CL_DYN_' + k.upper()] == v, k
    kinds = enum_members('cl_obs_kind')
    assert (kinds['CL_OBS_TS'], kinds['CL_OBS_DYN'], kinds['CL_OBS_OUTAGE'], kinds['CL_OBS_STATE']) == (S.OBS_TS, S.OBS_DYN, S.OBS_OUTAGE, S.OBS_STATE)
    assert int(re.search(r'#define CL_MAX_PHASES (\d+)', HEADER).group(1)) == 4 and S.CC_SLOTS == 6
    assert int(re.search(r'#define CL_MAX_CURVE (\d+)', HEADER).group(1)) == S.MAX_CURVE


def test_flag_bits_match_python():
    for name in ('HEATING_IS_HEAT_PUMP', 'DHW_IS_HEAT_PUMP', 'SIMULATE_OUTAGE', 'DYNAMICS', 'HAS_THERMAL', 'CS_HAS_MAX_IN',
                 'CS_HAS_MAX_OUT', 'HS_HAS_MAX_IN', 'HS_HAS_MAX_OUT', 'DS_HAS_MAX_IN', 'DS_HAS_MAX_OUT'):
        shift = int(re.search(r'#define CL_F_' + name + r'\s+\(1 << (\d+)\)', HEADER).group(1))
        assert getattr(S, 'F_' + name) == 1 << shift


def test_library_exports_every_declared_symbol():
    """`libcitylearn_b200.so` loads on a CPU-only box and exports every entry point the header declares."""
    from citylearn_b200 import _native, build
    build.build()
    lib = ctypes.CDLL(str(_native.library_path()))
    declared = re.findall(r'^\s*(?:int|const char\*)\s+(cl_\w+)\s*\(', HEADER, re.M)
    assert len(declared) >= 12
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.cl_abi_version.restype = ctypes.c_int
    assert lib.cl_abi_version() == _native.ABI_VERSION


def test_descriptor_struct_layout():
    from citylearn_b200._native import DistrictDesc, EvDesc
    # 12 int32 + 8 double + 6 pointers, no padding surprises (ABI 2: + the cl_ev_desc pointer)
    assert ctypes.sizeof(DistrictDesc) == 12 * 4 + 8 * 8 + 6 * 8
    assert DistrictDesc.ev.offset == 12 * 4 + 8 * 8 + 5 * 8
    # cl_ev_desc: 3 int32 (+ 4 bytes of padding before the first pointer) + 11 pointers + int32 (+ padding) + 4 pointers
    assert ctypes.sizeof(EvDesc) == 16 + 11 * 8 + 8 + 4 * 8 and EvDesc.ev_params.offset == 16 and EvDesc.cc_building.offset == 16 + 11 * 8 + 8
    assert int(re.search(r'#define CL_ABI_VERSION (\d+)', HEADER).group(1)) == 2


def test_charger_param_enum_matches_python():
    from citylearn_b200 import ev
    p = enum_members('cl_charger_param')
    assert p['CL_NCHP'] == len(ev.CHARGER_PARAMS)
    for k, v in ev.CHARGER_PARAMS.items():
        if 'CL_CH_' + k in p:
            assert p['CL_CH_' + k] == v, k
    assert {'CL_CH_MAX_C', 'CL_CH_EFF', 'CL_CH_C_N', 'CL_CH_D_N', 'CL_CH_C_X0', 'CL_CH_C_Y0', 'CL_CH_D_X0', 'CL_CH_D_Y0'} <= set(p)
    assert enum_members('cl_reward_id')['CL_REWARD_ELECTRIC_VEHICLES'] == 6


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from citylearn_b200 import CityLearnEnv
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        CityLearnEnv('citylearn_challenge_2022_phase_1')


def test_step_kernel_resource_budget():
    """Regression guard (no GPU: reads the cubin's resource table): the headline instantiations of the step kernel must stay spill-free
    and under the register count that lets one 512-thread block per SM be resident.  (Code generation of this kernel is fragile: a build
    with nvcc's --split-compile once put 8+ registers and 170-900 B of spills into every instantiation - 8 % of the C2 step time - without
    failing a single parity test.)"""
    import shutil
    import subprocess
    from citylearn_b200 import _native, build
    tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not Path(tool).exists():
        pytest.skip('cuobjdump not available')
    build.build()
    out = subprocess.run([tool, '-res-usage', str(_native.library_path())], capture_output=True, text=True, check=True).stdout
    usage = {}
    for name, reg, stack in re.findall(r'Function (\S*advance_kernel\S*):\s*\n\s*REG:(\d+) STACK:(\d+)', out):
        usage[name] = (int(reg), int(stack))
    assert len(usage) >= 20

    def find(real, thermal, dynamics, maxt, wide, kpi, ev, dec=0):
        b = lambda v: f'Lb{int(v)}E'  # noqa: E731
        key = f'advance_kernelI{real}{b(thermal)}{b(dynamics)}Li{maxt}E{b(wide)}{b(kpi)}{b(ev)}{b(dec)}E'
        hits = [v for k, v in usage.items() if key in k]
        assert len(hits) == 1, key
        return hits[0]
    # BASELINE configs[1] (2022 districts): fp64 flow and fp32, plain and with fused KPI accumulators
    for real, max_reg in (('d', 124), ('f', 112)):
        for dec in (0, 1):           # dec = 1: the opt-in split-phase-barrier instantiation
            reg, stack = find(real, 0, 0, 512, 0, 0, 0, dec)
            assert reg <= (128 if dec else max_reg) and stack == 0, (real, dec, reg, stack)
    assert find('d', 0, 0, 512, 0, 1, 0)[1] == 0 and find('f', 0, 0, 512, 0, 1, 0)[1] == 0
    # wide (building-tiled) districts, BASELINE configs[3]
    assert find('d', 0, 0, 512, 1, 0, 0)[1] <= 16 and find('f', 0, 0, 512, 1, 0, 0)[1] == 0
    # every 512-thread instantiation must fit one block per SM: 65536 registers / 512 threads
    for k, (reg, _) in usage.items():
        if 'Li512E' in k:
            assert reg <= 128, k
