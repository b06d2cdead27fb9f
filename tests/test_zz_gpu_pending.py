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
