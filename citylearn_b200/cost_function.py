"""KPI primitives (mirror of `citylearn/cost_function.py:10-388`) in plain NumPy.

Same names, arguments and return convention as the reference's `CostFunction` static methods: every function returns the
*running* series (one value per time step) as a list, the KPI of an episode being its last element.  The reference builds
them from pandas rolling windows of the full length; here they are cumulative sums / expanding means, which is what those
windows compute.  NaN handling follows pandas (`rolling(...).sum()` skips NaN, a window without valid values yields NaN).
"""
from __future__ import annotations

from typing import List, Tuple, Union

import numpy as np

DEFAULT_COMFORT_BAND = 2.0   # EnergySimulation.DEFUALT_COMFORT_BAND, citylearn/data.py:342


def _a(x) -> np.ndarray:
    return np.asarray(x, dtype='float64')


def _running_sum(x: np.ndarray) -> np.ndarray:
    """pandas `rolling(window=n, min_periods=1).sum()`: NaNs are skipped; NaN while no valid value has been seen."""
    valid = ~np.isnan(x)
    s = np.cumsum(np.where(valid, x, 0.0))
    return np.where(np.cumsum(valid) > 0, s, np.nan)


def _running_mean(x: np.ndarray) -> np.ndarray:
    valid = ~np.isnan(x)
    n = np.cumsum(valid)
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.where(n > 0, np.cumsum(np.where(valid, x, 0.0)) / n, np.nan)


def _groups(n: int, window: int) -> np.ndarray:
    return (np.arange(n) / window).astype(int)


class CostFunction:
    @staticmethod
    def ramping(net_electricity_consumption, down_ramp: bool = None, net_export: bool = None) -> List[float]:
        down_ramp = False if down_ramp is None else down_ramp
        net_export = True if net_export is None else net_export
        x = _a(net_electricity_consumption)
        r = np.full(x.shape, np.nan)
        r[1:] = x[1:] - x[:-1]
        r = np.abs(r) if down_ramp else np.where(np.isnan(r), np.nan, np.maximum(r, 0.0))
        if not net_export:
            r = np.where(x < 0, 0.0, r)
        return _running_sum(r).tolist()

    @staticmethod
    def one_minus_load_factor(net_electricity_consumption, window: int = None) -> List[float]:
        window = 730 if window is None else window
        x = _a(net_electricity_consumption)
        g = _groups(len(x), window)
        out = []
        for k in range(g.max() + 1 if len(x) else 0):
            v = x[g == k]
            with np.errstate(invalid='ignore', divide='ignore'):
                out.append(1 - np.nanmean(v) / np.nanmax(v))
        return _running_mean(np.array(out)).tolist()

    @staticmethod
    def peak(net_electricity_consumption, window: int = None) -> List[float]:
        window = 24 if window is None else window
        x = _a(net_electricity_consumption)
        g = _groups(len(x), window)
        out = np.array([np.nanmax(x[g == k]) for k in range(g.max() + 1 if len(x) else 0)])
        return _running_mean(out).tolist()

    @staticmethod
    def electricity_consumption(net_electricity_consumption) -> List[float]:
        return _running_sum(np.clip(_a(net_electricity_consumption), 0, None)).tolist()

    @staticmethod
    def zero_net_energy(net_electricity_consumption) -> List[float]:
        return _running_sum(_a(net_electricity_consumption)).tolist()

    @staticmethod
    def carbon_emissions(carbon_emissions) -> List[float]:
        return _running_sum(np.clip(_a(carbon_emissions), 0, None)).tolist()

    @staticmethod
    def cost(cost) -> List[float]:
        return _running_sum(np.clip(_a(cost), 0, None)).tolist()

    @staticmethod
    def quadratic(net_electricity_consumption) -> List[float]:
        return _running_sum(np.clip(_a(net_electricity_consumption), 0, None) ** 2).tolist()

    @staticmethod
    def discomfort(indoor_dry_bulb_temperature, dry_bulb_temperature_cooling_set_point, dry_bulb_temperature_heating_set_point,
                   band: Union[float, List[float]] = None, occupant_count=None) -> Tuple[list, ...]:
        t_in = _a(indoor_dry_bulb_temperature)
        n = len(t_in)
        occ = np.ones(n) if occupant_count is None else _a(occupant_count)
        band = np.broadcast_to(_a(DEFAULT_COMFORT_BAND if band is None else band), (n,))
        occupied = float(np.count_nonzero(occ > 0.0))
        cooling_delta = np.where(occ == 0.0, 0.0, t_in - _a(dry_bulb_temperature_cooling_set_point))
        heating_delta = np.where(occ == 0.0, 0.0, t_in - _a(dry_bulb_temperature_heating_set_point))
        with np.errstate(invalid='ignore', divide='ignore'):
            hot = (cooling_delta > band).astype('float64')
            cold = (heating_delta < -band).astype('float64')
            both = np.maximum(hot, cold)
            d = np.cumsum(both) / occupied
            dc = np.cumsum(cold) / occupied
            dh = np.cumsum(hot) / occupied
        cold_abs = np.abs(np.where(np.isnan(heating_delta), np.nan, np.minimum(heating_delta, 0.0)))
        hot_abs = np.abs(np.where(np.isnan(cooling_delta), np.nan, np.maximum(cooling_delta, 0.0)))

        def run(x, f):
            out = np.full(n, np.nan)
            acc = np.nan
            for i, v in enumerate(x):
                if not np.isnan(v):
                    acc = v if np.isnan(acc) else f(acc, v)
                out[i] = acc
            return out
        return (d.tolist(), dc.tolist(), dh.tolist(), run(cold_abs, min).tolist(), run(cold_abs, max).tolist(), _running_mean(cold_abs).tolist(),
                run(hot_abs, min).tolist(), run(hot_abs, max).tolist(), _running_mean(hot_abs).tolist())

    @staticmethod
    def one_minus_thermal_resilience(power_outage, **kwargs) -> List[float]:
        power_outage = np.array(power_outage, dtype='float32')
        occ = np.ones(len(power_outage), dtype='float32') if kwargs.get('occupant_count') is None else np.array(kwargs['occupant_count'], dtype='float32')
        occ[power_outage == 0.0] = 0.0
        kwargs['occupant_count'] = occ
        return CostFunction.discomfort(**kwargs)[0]

    @staticmethod
    def normalized_unserved_energy(expected_energy, served_energy, power_outage=None) -> List[float]:
        expected = _a(expected_energy).copy()
        unserved = expected - _a(served_energy)
        if power_outage is not None:
            off = _a(power_outage) == 0
            unserved[off] = 0.0
            expected[off] = 0.0
        with np.errstate(invalid='ignore', divide='ignore'):
            return (_running_sum(unserved) / np.nansum(expected)).tolist()
