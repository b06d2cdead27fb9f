"""Online KPI accumulators (cl_kpi_*): the batched KPI ratios must equal the history-based `evaluate()` table, which
tests/test_evaluate.py pins against the reference's own evaluate()."""
import numpy as np
import pytest

from citylearn_b200 import schema as S
from citylearn_b200.evaluate import KE, KU, History, _push, evaluate, evaluate_batched
from citylearn_oracle import OracleEnv

DISTRICT_KPIS = ['electricity_consumption_total', 'zero_net_energy', 'carbon_emissions_total', 'cost_total', 'ramping_average',
                 'daily_one_minus_load_factor_average', 'monthly_one_minus_load_factor_average', 'daily_peak_average', 'all_time_peak_average']


def accumulate_numpy(spec, dyn, district, start):
    """What kpi_accumulate_kernel does, step by step, for ONE env: dyn [k, B, NDYN] float32, district [k, 3] float32."""
    B = spec.n_buildings
    unit = np.zeros((1, B, 8))
    env = np.zeros((2, 1, len(KE)))
    D = S.DYN
    for k in range(dyn.shape[0]):
        net = dyn[k, :, D['net_electricity_consumption']].astype('float64')
        sto = sum(dyn[k, :, D[f'{n}_storage_electricity_consumption']].astype('float64') for n in ('cooling', 'heating', 'dhw', 'electrical'))
        nws = net - sto
        price = np.array([b.series['electricity_pricing'][start + k] for b in spec.buildings], dtype='float32').astype('float64')
        carbon = np.array([b.series['carbon_intensity'][start + k] for b in spec.buildings], dtype='float32').astype('float64')
        unit[0, :, KU['ec']] += np.maximum(net, 0); unit[0, :, KU['zne']] += net
        unit[0, :, KU['emission']] += np.maximum(dyn[k, :, D['net_electricity_consumption_emission']].astype('float64'), 0)
        unit[0, :, KU['cost']] += np.maximum(dyn[k, :, D['net_electricity_consumption_cost']].astype('float64'), 0)
        unit[0, :, KU['b_ec']] += np.maximum(nws, 0); unit[0, :, KU['b_zne']] += nws
        unit[0, :, KU['b_emission']] += np.maximum(carbon * nws, 0); unit[0, :, KU['b_cost']] += np.maximum(price * nws, 0)
        env[0] = _push(env[0], float(district[k, 0]))
        env[1] = _push(env[1], float(nws.sum()))
    return unit, np.stack([env[0][0], env[1][0]])[None]


def history_table(spec, dyn, district, start, T):
    h = History(dyn, district, start, np.zeros((spec.n_buildings, T), dtype='float32'))
    rows = evaluate(spec, h, as_dataframe=False)
    return {r['cost_function']: r['value'] for r in rows if r['level'] == 'district'}, \
        {(r['name'], r['cost_function']): r['value'] for r in rows if r['level'] == 'building'}


@pytest.mark.parametrize('dataset,steps', [('citylearn_challenge_2022_phase_1', 100), ('citylearn_challenge_2022_phase_1', 791),
                                           ('citylearn_challenge_2020_climate_zone_1', 60)])
def test_batched_finalisation_matches_history_evaluate(dataset, steps):
    spec = S.load(dataset)
    env = OracleEnv(spec, 1)
    env.reset()
    rng = np.random.RandomState(4)
    dyn, dist = [], []
    for k in range(steps):
        _, _, d, y = env.step(rng.uniform(-1, 1, size=(1, spec.action_dim)).astype('float32'))
        dyn.append(y[0].astype('float32')); dist.append(d[0].astype('float32'))
    dyn, dist = np.stack(dyn), np.stack(dist)
    unit, envacc = accumulate_numpy(spec, dyn, dist, spec.simulation_start_time_step)
    got = evaluate_batched(spec, unit, envacc)
    ref_d, ref_b = history_table(spec, dyn, dist, spec.simulation_start_time_step, spec.simulation_end_time_step - spec.simulation_start_time_step + 1)
    for name in DISTRICT_KPIS:
        assert got['district'][name][0] == pytest.approx(ref_d[name], rel=1e-9, abs=1e-12), name
    for bi, b in enumerate(spec.buildings):
        for name in DISTRICT_KPIS[:4]:
            v = ref_b[(b.name, name)]
            g = got['building'][name][0, bi]
            assert (np.isnan(g) and v is None) or g == pytest.approx(v, rel=1e-9, abs=1e-12), (b.name, name)


@pytest.mark.gpu
def test_device_accumulators_match_history_evaluate_for_every_env():
    """E envs with different action sequences: device accumulators -> batched KPIs == history-based evaluate() of each env."""
    import torch
    from citylearn_b200 import CityLearnEnv
    E, K = 6, 300
    spec = S.load('citylearn_challenge_2022_phase_1')
    env = CityLearnEnv(spec, num_envs=E, track_kpis=True, record_history=True, history_env=2)
    oracle = OracleEnv(spec, E)
    env.reset(); oracle.reset()
    rng = np.random.RandomState(8)
    dyn, dist = [], []
    for k in range(K):
        a = rng.uniform(-1, 1, size=(E, spec.action_dim)).astype('float32')
        env.step(torch.from_numpy(a).cuda())
        _, _, d, y = oracle.step(a)
        dyn.append(y.astype('float32')); dist.append(d.astype('float32'))
    got = env.evaluate_batched()
    dyn, dist = np.stack(dyn), np.stack(dist)            # [K, E, B, NDYN], [K, E, 3]
    T = env.time_steps
    for e in range(E):
        ref_d, ref_b = history_table(spec, dyn[:, e], dist[:, e], spec.simulation_start_time_step, T)
        for name in DISTRICT_KPIS:
            assert got['district'][name][e] == pytest.approx(ref_d[name], rel=2e-6, abs=1e-9), (e, name)
    # and the recorded env's own table agrees with its batched row
    df = env.evaluate()
    table = {r['cost_function']: r['value'] for r in df.to_dict('records') if r['level'] == 'district'}
    for name in DISTRICT_KPIS:
        assert got['district'][name][2] == pytest.approx(table[name], rel=2e-6, abs=1e-9), name
    with pytest.raises(RuntimeError):
        CityLearnEnv(spec, num_envs=2).evaluate_batched()


@pytest.mark.gpu
def test_fused_accumulators_survive_rollouts_and_equal_the_trace_fed_kernel(monkeypatch):
    """The accumulators live inside advance_kernel: a mix of step(), rollout() and CUDA-graph replays gives the KPIs of the
    trace-fed second kernel (CL_B200_KPI_UNFUSED, the round-1 path) stepping one by one - and no trace buffer is allocated."""
    import torch
    from citylearn_b200 import CityLearnEnv
    from citylearn_b200.closed_loop import ClosedLoop
    E, K = 40, 96
    spec = S.load('citylearn_challenge_2022_phase_1')
    g = torch.Generator(device='cuda').manual_seed(5)
    acts = torch.rand((K, E, spec.action_dim), device='cuda', generator=g) * 2 - 1
    fused = CityLearnEnv(spec, num_envs=E, track_kpis=True)
    assert fused._kpi_fused and fused.trace is None
    fused.reset()
    for k in range(10):
        fused.step(acts[k])
    fused.rollout(acts[10:50].contiguous(), None, torch.empty((40, E, spec.n_buildings), device='cuda'), None)     # one 40-step launch
    idx = torch.zeros(1, dtype=torch.long, device='cuda')          # step counter on the device: the captured policy reads acts[idx]

    def pol(obs):
        a = acts.index_select(0, idx.clamp(max=K - 1))[0]
        idx.add_(1)
        return a
    loop = ClosedLoop(fused, pol, steps_per_replay=8)
    idx.fill_(50)
    loop.run(K - 50)                                               # 5 graph replays of 8 steps + 6 single-step replays
    got = fused.evaluate_batched()
    monkeypatch.setenv('CL_B200_KPI_UNFUSED', '1')
    ref = CityLearnEnv(spec, num_envs=E, track_kpis=True)
    assert not ref._kpi_fused and ref.trace is not None
    ref.reset()
    for k in range(K):
        ref.step(acts[k])
    exp = ref.evaluate_batched()
    for level in ('district', 'building'):
        for name, v in exp[level].items():
            np.testing.assert_allclose(got[level][name], v, rtol=1e-12, atol=0, equal_nan=True, err_msg=f'{level} {name}')
