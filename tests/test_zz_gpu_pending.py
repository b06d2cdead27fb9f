"""GPU tests written after the round's GPU budget was spent: pinned on the CPU (oracle vs the reference), not yet run on hardware.
Non-strict xfail and last in collection order, so that nothing here can affect the validated suite."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason='not yet run on hardware')
def test_dual_mode_cooling_or_heating_device_matches_reference():
    """`cooling_or_heating_device` (one signed action for both heat pumps), hvac modes 2 / 3, heating heat pump under LSTM dynamics
    (citylearn_b200.synthetic.SyntheticDualModeSource).  The oracle reproduces the reference's trace (tests/test_oracle_golden.py)."""
    import test_gpu_parity as G
    G.LSTM_CASES.append('c9_dual_mode')
    try:
        G.test_single_env_matches_reference_traces('c9_dual_mode')
    finally:
        G.LSTM_CASES.remove('c9_dual_mode')


@pytest.mark.xfail(strict=False, reason='not yet run on hardware')
@pytest.mark.parametrize('fixture', ['trace_fuzz.json.gz', 'trace_datasets.json.gz'])
def test_fuzzed_reference_runs_on_gpu(fixture):
    """The short reference runs of tests/test_oracle_golden.py::test_oracle_matches_fuzzed_reference_runs through the C ABI."""
    import gzip
    import json
    import numpy as np
    from citylearn_b200 import CityLearnEnv
    from helpers import GOLDEN, max_abs_diff, schema_for
    for c in json.load(gzip.open(GOLDEN / fixture, 'rt'))['cases']:
        sch, src, ov = schema_for({'dataset': c['dataset'], 'reward': c['reward'], 'overrides': c['overrides']})
        env = CityLearnEnv(sch, data_source=src, num_envs=1, **ov)
        lstm = any(b.dynamics for b in env.spec.buildings)
        tag = (c['dataset'], c['overrides'])
        obs, _ = env.reset()
        assert max_abs_diff(np.array([v for row in obs for v in row], dtype='float32'), np.array(c['reset_obs'], dtype='float32')) == 0.0, tag
        acts = np.array(c['actions'], dtype='float32')
        for k in range(len(acts)):
            obs, rew, _, _, _ = env.step(acts[k][None])
            assert max_abs_diff(obs.cpu().numpy()[0], np.array(c['obs'][k], dtype='float32')) == 0.0, (tag, k)
            ref_r = np.array(c['reward_values'][k], dtype='float32')
            assert max_abs_diff(rew.cpu().numpy()[0], ref_r) <= 1e-5 * max(1.0, float(np.abs(ref_r).max())), (tag, k)
            assert max_abs_diff(env.district.cpu().numpy()[0], np.array(c['district'][k], dtype='float32')) <= (1e-6 if lstm else 0.0) * 1.0 \
                or max_abs_diff(env.district.cpu().numpy()[0], np.array(c['district'][k], dtype='float32')) <= 2.4e-7 * float(np.abs(c['district'][k]).max()), (tag, k)
