"""The exact per-unit device code (citylearn_b200/csrc/unit_physics.cuh), compiled for the host with g++, vs the oracle.

CPU-side coverage of the kernel physics; the GPU parity tests proper are in test_gpu_parity.py (-m gpu).
"""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from citylearn_b200 import schema as S
from citylearn_b200.schema import DYN, P
from citylearn_oracle import OracleEnv
from helpers import actions_of, load_golden, spec_for, within_scaled_tolerance

HERE = Path(__file__).resolve().parent / 'host'


@pytest.fixture(scope='module')
def hostlib():
    so = HERE / 'libhost_physics.so'
    src = HERE / 'host_physics.cpp'
    deps = [src, HERE.parents[1] / 'citylearn_b200' / 'csrc' / 'unit_physics.cuh', HERE.parents[1] / 'include' / 'citylearn_b200.h']
    if not so.is_file() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC', '-o', str(so), str(src)], check=True)
    return ctypes.CDLL(str(so))


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_host(hostlib, spec, acts, precision, E=1):
    env = OracleEnv(spec, E)
    env.reset()
    B = spec.n_buildings
    U = B * E
    Pm = np.ascontiguousarray(spec.params.T, dtype='float64')
    ipm = np.ascontiguousarray(spec.iparams.T, dtype='int32')
    table = np.ascontiguousarray(spec.table)
    start = np.full(E, spec.simulation_start_time_step, dtype='int32')
    state = np.zeros((6, U), dtype='float64')
    tile = lambda v: np.tile(v, E)   # noqa: E731
    state[0] = tile(np.float32(spec.params[:, P['BAT_INITIAL_SOC']]))
    state[1] = tile(spec.params[:, P['BAT_CAPACITY']])
    state[2] = tile(np.sqrt(spec.params[:, P['BAT_EFFICIENCY0']]))
    state[3] = tile(np.float32(spec.params[:, P['CS_INITIAL_SOC']]))
    state[4] = tile(np.float32(spec.params[:, P['HS_INITIAL_SOC']]))
    state[5] = tile(np.float32(spec.params[:, P['DS_INITIAL_SOC']]))
    outage = np.ascontiguousarray(env.outage, dtype='float32')
    dyn = np.zeros((U, S.NDYN), dtype='float32')
    out = []
    for k in range(len(acts)):
        a = np.ascontiguousarray(acts[k].reshape(E, -1), dtype='float32')
        ctrl = None
        if env.has_dyn:
            ctrl = np.full(U, 1 if env.window_fill > env.L_max() else 0, dtype='uint8')
        _, _, _, odyn = env.step(a)
        hostlib.host_step(precision, B, E, table.shape[1], ptr(Pm), ptr(ipm), ptr(table), ptr(start), k, ptr(outage), env.T, ptr(a),
                          a.shape[1], ptr(state), ptr(dyn), ptr(ctrl) if ctrl is not None else None)
        out.append((dyn.reshape(E, B, -1).copy(), odyn))
    return out


CHECKED = ['electrical_storage_soc', 'electrical_storage_energy_balance', 'electrical_storage_electricity_consumption',
           'net_electricity_consumption', 'net_electricity_consumption_cost', 'net_electricity_consumption_emission',
           'cooling_storage_soc', 'dhw_storage_soc', 'dhw_storage_energy_balance', 'cooling_electricity_consumption',
           'heating_electricity_consumption', 'dhw_electricity_consumption', 'cooling_demand', 'dhw_demand',
           'non_shiftable_load_electricity_consumption', 'heating_storage_soc', 'heating_storage_energy_balance', 'heating_demand',
           'cooling_storage_energy_balance', 'dhw_storage_soc']


@pytest.mark.parametrize('case', ['c1_phase1_300', 'c1_subhour', 'c2_marl', 'c3_marl', 'c6_tanks_2020_marl_central', 'c6_baeda3',
                                  'c8_heating', 'c8_heating_central_marl'])
def test_fp64_flow_is_bit_exact(hostlib, case):
    z, cfg, _ = load_golden(case)
    spec = spec_for(cfg)
    for got, ref in run_host(hostlib, spec, actions_of(z)[0], precision=1):
        for n in CHECKED:
            assert np.array_equal(got[..., DYN[n]], ref[..., DYN[n]].astype('float32'), equal_nan=True), n


@pytest.mark.parametrize('case', ['c1_phase1_300', 'c3_marl'])
def test_fp32_flow_is_close(hostlib, case):
    """Plain float arithmetic: the battery's steep capacity-power curve amplifies rounding, so this mode is only held to
    1e-4 scaled-relative here (it is NOT the default precision; DESIGN.md 'Numerics')."""
    z, cfg, _ = load_golden(case)
    spec = spec_for(cfg)
    worst = 0.0
    for got, ref in run_host(hostlib, spec, actions_of(z)[0], precision=0):
        for n in CHECKED:
            scale = 1.0 if n.endswith('_soc') else 10.0
            ok, w = within_scaled_tolerance(got[..., DYN[n]], ref[..., DYN[n]], scale, rtol=1e-4)
            worst = max(worst, w)
            assert ok, (n, w)
    assert worst > 0.0


@pytest.mark.parametrize('dataset', ['citylearn_challenge_2022_phase_1', 'citylearn_challenge_2020_climate_zone_1'])
@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_fp64_flow_fuzzed_parameters(hostlib, dataset, seed):
    """Randomly perturbed device parameters (capacities, powers, efficiencies, losses, depth of discharge, sub-hour ratios) and 4 envs with extreme / zero / saturating actions: the host-compiled device code and the oracle must
    stay bit-identical in regimes no dataset exercises."""
    import copy
    spec = copy.deepcopy(S.load(dataset))
    rng = np.random.RandomState(100 + seed)
    B = spec.n_buildings
    p = spec.params
    u = lambda lo, hi: rng.uniform(lo, hi, B)          # noqa: E731
    p[:, P['BAT_CAPACITY']] *= u(0.2, 3.0)
    p[:, P['BAT_NOMINAL_POWER']] *= u(0.3, 2.0)
    p[:, P['BAT_LOSS']] = u(0.0, 0.01)
    p[:, P['BAT_CLC']] = u(0.0, 1e-3)
    p[:, P['BAT_DOD']] = u(0.6, 1.0)
    p[:, P['BAT_INITIAL_SOC']] = np.float32(u(0.0, 0.5))
    if seed % 2:
        p[:, P['BAT_CAPACITY']][0] = 0.0               # an absent battery among real ones
    thermal = bool((spec.iparams[:, S.IP['FLAGS']] & S.F_HAS_THERMAL).any())
    for pre in (('CS', 'HS', 'DS') if thermal else ()):
        p[:, P[f'{pre}_CAPACITY']] = np.float32(p[:, P[f'{pre}_CAPACITY']] * u(0.3, 2.0))
        p[:, P[f'{pre}_EFFICIENCY']] = u(0.8, 1.0)
        p[:, P[f'{pre}_LOSS']] = u(0.0, 0.02)
        p[:, P[f'{pre}_INITIAL_SOC']] = np.float32(u(0.0, 0.9))
    for pre in ('CD', 'HD', 'DD'):
        p[:, P[f'{pre}_NOMINAL_POWER']] = np.float32(p[:, P[f'{pre}_NOMINAL_POWER']] * u(0.5, 1.5))
    if seed >= 2:                                      # sub-hour control steps on an hourly dataset
        p[:, P['TIME_STEP_RATIO']] = 0.5
        p[:, P['HOURS_PER_STEP']] = 0.5
    E, K = 4, 48
    acts = rng.uniform(-1, 1, size=(K, E, spec.action_dim)).astype('float32')
    acts[:, 1] = np.sign(acts[:, 1])                   # saturating
    acts[::3, 2] = 0.0                                 # idle steps
    acts[:, 3] *= 0.05                                 # tiny
    for got, ref in run_host(hostlib, spec, acts, precision=1, E=E):
        for n in CHECKED + ['electrical_storage_degraded_capacity']:
            a, b = got[..., DYN[n]], ref[..., DYN[n]].astype('float32')
            assert np.array_equal(a, b, equal_nan=True), (n, float(np.nanmax(np.abs(a.astype('float64') - b))))


@pytest.mark.parametrize('precision', [0, 1])
def test_curve_grid_search_matches_reference_scan(hostlib, precision):
    """The device-side uniform-grid segment search (SmemCurves + its index) picks the segment the reference's
    `argmax(x <= xs) - 1` picks (energy_model.py:1083-1109): dataset curves and random ascending curves, probed at random points,
    at / next to every abscissa and cell boundary, below 0, above 1, NaN and inf."""
    rng = np.random.RandomState(5)
    curves = [np.array([0.0, 0.3, 0.7, 0.8, 1.0]), np.array([0.0, 0.8, 1.0]), np.array([0.0, 1.0])]
    for _ in range(40):
        n = rng.randint(2, 9)
        cells = np.sort(rng.choice(33, size=n, replace=False))
        curves.append(np.minimum((cells + rng.uniform(0, 0.999, n)) / 32.0, 1.0 + 0 * cells))
    hostlib.host_curve_check.restype = ctypes.c_int
    n_indexed = 0
    for xs in curves:
        xs = np.ascontiguousarray(xs, dtype='float64')
        ys = np.ascontiguousarray(rng.uniform(0.1, 1.0, len(xs)))
        probes = [rng.uniform(-0.1, 1.2, 4000), xs, np.nextafter(xs, -1), np.nextafter(xs, 2), np.float32(xs).astype('float64'),
                  np.nextafter(np.float32(xs), np.float32(-1)).astype('float64'), np.nextafter(np.float32(xs), np.float32(2)).astype('float64'),
                  np.arange(34) / 32.0, np.nextafter(np.arange(34) / 32.0, -1), [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e300, 5e-324]]
        x = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype='float64') for p in probes]))
        first = ctypes.c_int(-1)
        bad = hostlib.host_curve_check(precision, ptr(xs), ptr(ys), len(xs), ptr(x), len(x), ctypes.byref(first))
        if bad == -1:
            continue
        n_indexed += 1
        assert bad == 0, f'curve {xs}: {bad} mismatches, first at x = {x[first.value]!r}'
    assert n_indexed >= 20


def test_charger_step_device_code_matches_the_oracle(hostlib):
    """`charger_step` + the vehicle battery's `battery_charge` (the code the EV instantiation of the step kernel runs), compiled for the
    host, against the oracle's charger update on the 300 reference actions of `c10_evs`: SOC entry, degraded capacity, the charger's
    electricity consumption and the commanded energy bit for bit (fp64 flow), every charger update of the run."""
    import json
    from citylearn_b200.data import DataSet
    from citylearn_b200.ev import CHARGER_PARAMS as CP
    from helpers import GOLDEN
    z = np.load(GOLDEN / 'ev' / 'c10_evs.npz')
    cfg = json.loads(bytes(z['config']).decode())
    src = DataSet.get_source(cfg['dataset'])
    sch = src.schema()
    sch['reward_function'] = {'type': cfg['reward']['type'], 'attributes': {}}
    spec = S.load(sch, data_source=src, ev_random_seed=cfg['np_seed'])
    ev = spec.ev
    env = OracleEnv(spec, 1)                       # sqrt (not libm pow), like the device code
    env.reset()
    hostlib.host_charger_step.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    updates = 0
    for k in range(len(z['actions'])):
        t = env.t
        row = int(env.start[0]) + t
        pre = []
        for c_i in range(len(ev['chargers'])):
            conn = spec.table[row, ev['ch_cols'][c_i, 0]] > 0
            v = max(int(spec.table[row, ev['ch_cols'][c_i, 1]]), 0)
            soc = env.ev_soc_prev[0, v] if t > 0 else env.ev_soc[0, v]
            pre.append((conn, v, float(np.float32(soc)), float(env.ev_cap_deg[0, v]), float(np.sqrt(env.ev_eff[0, v])), bool(env.ev_charged[0, v])))
        env.step(z['actions'][k][None])
        info = env.last_ev
        for c_i, (conn, v, soc, cap, rte, was_charged) in enumerate(pre):
            slot = int(ev['ch_action'][c_i])
            act = float(z['actions'][k][slot]) if slot >= 0 else 0.0
            state = np.array([soc, cap, rte], dtype='float64')
            out = np.zeros(2, dtype='float64')
            chp = np.ascontiguousarray(ev['ch_params'][c_i], dtype='float64')
            evp = np.ascontiguousarray(ev['ev_params'][v], dtype='float64')
            charged = hostlib.host_charger_step(1, ptr(chp), act, int(conn), ptr(evp), int(ev['ev_ip'][v, 0]), int(ev['ev_ip'][v, 1]), int(not was_charged),
                                                ptr(state), ptr(out))
            assert np.float32(out[0]) == info['ch_ec'][0, c_i], (k, c_i)
            assert np.float32(out[1]) == info['past'][0, c_i], (k, c_i)
            if charged:
                updates += 1
                assert np.float32(state[0]) == np.float32(env.ev_soc_prev[0, v]), (k, c_i)       # soc[t], one next_time_step later
                assert state[1] == env.ev_cap_deg[0, v] and state[2] == np.sqrt(env.ev_eff[0, v]), (k, c_i)
    assert updates > 300
