"""`CityLearnEnv.evaluate()` - end-of-episode KPI table (mirror of `citylearn/citylearn.py:1136-1323`).

Works on the per-step history of ONE environment (`History`): the per-building values the step kernel reports in its trace
(`cl_dyn` slots) plus the district sums.  The reference evaluates slices `[0 : time_step + 1]` of its per-building arrays -
whose last entry has not been simulated yet (zero-initialised / dataset values) - against district lists that have only
`time_step` entries; both quirks are reproduced because they change the ramping / load-factor / peak ratios.

Returns a list of records (`cost_function`, `value`, `name`, `level`) or, when pandas is importable, the same as a DataFrame
with the reference's column order.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import schema as S
from .cost_function import CostFunction, DEFAULT_COMFORT_BAND

CONDITIONS = {
    'WITH_STORAGE_AND_PV': '', 'WITHOUT_STORAGE_BUT_WITH_PV': '_without_storage', 'WITHOUT_STORAGE_AND_PV': '_without_storage_and_pv',
    'WITH_STORAGE_AND_PARTIAL_LOAD_AND_PV': '', 'WITHOUT_STORAGE_BUT_WITH_PARTIAL_LOAD_AND_PV': '_without_storage',
    'WITHOUT_STORAGE_AND_PARTIAL_LOAD_BUT_WITH_PV': '_without_storage_and_partial_load',
    'WITHOUT_STORAGE_AND_PARTIAL_LOAD_AND_PV': '_without_storage_and_partial_load_and_pv',
}


@dataclass
class History:
    """Values of one env for steps 0 .. k-1: `dyn` [k, B, NDYN] (float32), `district` [k, 3], window start row, outage [B, T]."""
    dyn: np.ndarray
    district: np.ndarray
    episode_start: int
    outage: np.ndarray


def _safe_div(c, b) -> Optional[float]:
    def coerce(x):
        try:
            v = float(x)
            return v if np.isfinite(v) else 0.0
        except Exception:
            return 0.0
    c, b = coerce(c), coerce(b)
    if b == 0.0:
        return 1.0 if c == 0.0 else None
    return c / b


def _suffix(condition) -> str:
    if condition is None:
        return None
    if isinstance(condition, str):
        return CONDITIONS.get(condition, condition)
    return getattr(condition, 'value', condition)      # a reference `EvaluationCondition` enum member


class _BuildingSeries:
    """The reference's per-building series properties (`citylearn/building.py:307-637, 2615-2703, 2886-2932`) on `[0 : k + 1]`."""

    def __init__(self, spec: S.DistrictSpec, bi: int, h: History):
        b = spec.buildings[bi]
        self.b = b
        k = h.dyn.shape[0]
        n = k + 1
        rows = slice(h.episode_start, h.episode_start + n)
        s = b.series
        D = lambda name: h.dyn[:, bi, S.DYN[name]].astype('float64')          # noqa: E731
        ext = lambda a, last: np.concatenate([a, [last]])                     # noqa: E731
        self.n = n
        self.net = ext(D('net_electricity_consumption'), 0.0)
        self.cost = ext(D('net_electricity_consumption_cost'), 0.0)
        self.emission = ext(D('net_electricity_consumption_emission'), 0.0)
        storage = sum(ext(D(f'{x}_storage_electricity_consumption'), 0.0) for x in ('cooling', 'heating', 'dhw', 'electrical'))
        self.price = s['electricity_pricing'][rows].astype('float64')
        self.carbon = s['carbon_intensity'][rows].astype('float64')
        self.solar = (S.pv_generation(b, s['solar_generation'][rows]) * -1)
        self.t_out = s['outdoor_dry_bulb_temperature'][rows]
        self.net_wo_storage = self.net - storage
        # partial load (DynamicsBuilding.net_electricity_consumption_without_storage_and_partial_load, building.py:2909-2922)
        self.cool_dem = ext(D('cooling_demand_series'), float(s['cooling_demand'][rows][-1]))
        self.heat_dem = ext(D('heating_demand_series'), float(s['heating_demand'][rows][-1]))
        if b.dynamics:
            dv = b.devices
            cd = s['cooling_demand'][rows].astype('float32') - self.cool_dem.astype('float32')
            hd = s['heating_demand'][rows].astype('float32') - self.heat_dem.astype('float32')
            cool_diff = cd / S.cop32(dv['cooling_device'], self.t_out, False)
            if dv['heating_device']['type'] == 'HeatPump':
                heat_diff = hd / S.cop32(dv['heating_device'], np.array(self.t_out[-1]), True)     # scalar temperature quirk (:2917)
            else:
                heat_diff = hd / dv['dhw_device']['efficiency']                                      # dhw device quirk (:2919)
            self.net_wo_storage_partial = self.net_wo_storage + (cool_diff.astype('float64') + np.asarray(heat_diff, dtype='float64'))
        else:
            self.net_wo_storage_partial = self.net_wo_storage
        # comfort / resilience inputs
        self.t_in = ext(D('indoor_dry_bulb_temperature'), float(s['indoor_dry_bulb_temperature'][rows][-1]))
        self.cool_sp = s['indoor_dry_bulb_temperature_cooling_set_point'][rows].astype('float64')
        self.heat_sp = s['indoor_dry_bulb_temperature_heating_set_point'][rows].astype('float64')
        self.occupants = s['occupant_count'][rows].astype('float64')
        self.outage = np.asarray(h.outage[bi, :n], dtype='float64')
        dhw, nsl = s['dhw_demand'][rows].astype('float64'), s['non_shiftable_load'][rows].astype('float64')
        self.expected = self.cool_dem + self.heat_dem + dhw + nsl
        self.served = (ext(D('cooling_demand'), float(s['cooling_demand'][rows][-1])) + ext(D('heating_demand'), float(s['heating_demand'][rows][-1]))
                       + ext(D('dhw_demand'), float(dhw[-1])) + ext(D('energy_to_non_shiftable_load'), float(nsl[-1])))

    def net_series(self, suffix: str) -> np.ndarray:
        base = {'': self.net, '_without_storage': self.net_wo_storage, '_without_storage_and_pv': self.net_wo_storage - self.solar,
                '_without_storage_and_partial_load': self.net_wo_storage_partial,
                '_without_storage_and_partial_load_and_pv': self.net_wo_storage_partial - self.solar}
        return base[suffix]

    def cost_series(self, suffix: str) -> np.ndarray:
        return self.cost if suffix == '' else self.price * self.net_series(suffix)

    def emission_series(self, suffix: str) -> np.ndarray:
        return self.emission if suffix == '' else np.clip(self.carbon * self.net_series(suffix), 0, None)


def evaluate(spec: S.DistrictSpec, history: History, control_condition=None, baseline_condition=None, comfort_band: float = None,
             as_dataframe: bool = True):
    k = history.dyn.shape[0]
    if k < 1:
        raise RuntimeError('evaluate() needs at least one simulated step')
    comfort_band = DEFAULT_COMFORT_BAND if comfort_band is None else comfort_band
    control, baseline = _suffix(control_condition), _suffix(baseline_condition)
    rows: List[Dict] = []
    series: List[_BuildingSeries] = []
    for bi, b in enumerate(spec.buildings):
        # defaults are fixed by the first building's type (citylearn.py:1166-1177)
        if control is None:
            control = ''
        if baseline is None:
            baseline = '_without_storage_and_partial_load' if b.dynamics else '_without_storage'
        bs = _BuildingSeries(spec, bi, history)
        series.append(bs)
        kw = dict(indoor_dry_bulb_temperature=bs.t_in, dry_bulb_temperature_cooling_set_point=bs.cool_sp,
                  dry_bulb_temperature_heating_set_point=bs.heat_sp, band=comfort_band, occupant_count=bs.occupants)
        unmet, cold, hot, cmin, cmax, cavg, hmin, hmax, havg = CostFunction.discomfort(**kw)
        ec_c = CostFunction.electricity_consumption(bs.net_series(control))[-1]
        ec_b = CostFunction.electricity_consumption(bs.net_series(baseline))[-1]
        zne_c = CostFunction.zero_net_energy(bs.net_series(control))[-1]
        zne_b = CostFunction.zero_net_energy(bs.net_series(baseline))[-1]
        ce_c = CostFunction.carbon_emissions(bs.emission_series(control))[-1]
        ce_b = CostFunction.carbon_emissions(bs.emission_series(baseline))[-1] if float(b.series['carbon_intensity'].sum()) != 0 else 0
        co_c = CostFunction.cost(bs.cost_series(control))[-1]
        co_b = CostFunction.cost(bs.cost_series(baseline))[-1] if float(b.series['electricity_pricing'].sum()) != 0 else 0
        values = [
            ('electricity_consumption_total', _safe_div(ec_c, ec_b)), ('zero_net_energy', _safe_div(zne_c, zne_b)),
            ('carbon_emissions_total', _safe_div(ce_c, ce_b)), ('cost_total', _safe_div(co_c, co_b)),
            ('discomfort_proportion', unmet[-1]), ('discomfort_cold_proportion', cold[-1]), ('discomfort_hot_proportion', hot[-1]),
            ('discomfort_cold_delta_minimum', cmin[-1]), ('discomfort_cold_delta_maximum', cmax[-1]), ('discomfort_cold_delta_average', cavg[-1]),
            ('discomfort_hot_delta_minimum', hmin[-1]), ('discomfort_hot_delta_maximum', hmax[-1]), ('discomfort_hot_delta_average', havg[-1]),
            ('one_minus_thermal_resilience_proportion', CostFunction.one_minus_thermal_resilience(power_outage=bs.outage, **kw)[-1]),
            ('power_outage_normalized_unserved_energy_total', CostFunction.normalized_unserved_energy(bs.expected, bs.served, power_outage=bs.outage)[-1]),
            ('annual_normalized_unserved_energy_total', CostFunction.normalized_unserved_energy(bs.expected, bs.served)[-1]),
        ]
        rows += [{'cost_function': n, 'value': v, 'name': b.name, 'level': 'building'} for n, v in values]

    def district_series(suffix: str) -> np.ndarray:
        if suffix == '':
            return history.district[:, 0].astype('float64')       # the env's own list: one entry per simulated step
        return np.sum([s.net_series(suffix) for s in series], axis=0)
    dc, db = district_series(control), district_series(baseline)
    T = history.outage.shape[1]
    district = [
        ('ramping_average', _safe_div(CostFunction.ramping(dc)[-1], CostFunction.ramping(db)[-1])),
        ('daily_one_minus_load_factor_average', _safe_div(CostFunction.one_minus_load_factor(dc, window=24)[-1], CostFunction.one_minus_load_factor(db, window=24)[-1])),
        ('monthly_one_minus_load_factor_average', _safe_div(CostFunction.one_minus_load_factor(dc, window=730)[-1], CostFunction.one_minus_load_factor(db, window=730)[-1])),
        ('daily_peak_average', _safe_div(CostFunction.peak(dc, window=24)[-1], CostFunction.peak(db, window=24)[-1])),
        ('all_time_peak_average', _safe_div(CostFunction.peak(dc, window=T)[-1], CostFunction.peak(db, window=T)[-1])),
    ]
    # district table = mean over the district-level and building-level rows per cost function (citylearn.py:1310-1317)
    acc: Dict[str, List[float]] = {}
    for n, v in district + [(r['cost_function'], r['value']) for r in rows]:
        acc.setdefault(n, [])
        if v is not None and not (isinstance(v, float) and np.isnan(v)):
            acc[n].append(float(v))
    district_rows = [{'cost_function': n, 'value': (float(np.mean(v)) if len(v) else float('nan')), 'name': 'District', 'level': 'district'}
                     for n, v in sorted(acc.items())]
    records = district_rows + rows
    if as_dataframe:
        try:
            import pandas as pd
            return pd.DataFrame(records, columns=['cost_function', 'value', 'name', 'level'])
        except Exception:   # pragma: no cover
            pass
    return records


# ------------------------------------------------------------------------------------------------------------------
# batched envs: KPIs from the on-device accumulators (cl_kpi_*, include/citylearn_b200.h) instead of a per-step history
# ------------------------------------------------------------------------------------------------------------------
KU = {'ec': 0, 'zne': 1, 'emission': 2, 'cost': 3, 'b_ec': 4, 'b_zne': 5, 'b_emission': 6, 'b_cost': 7}
KE = {n: i for i, n in enumerate(['n', 'prev', 'ramp', 'all_max', 'd_sum', 'd_max', 'd_cnt', 'd_fin_lf', 'd_fin_peak', 'd_fin_n',
                                  'm_sum', 'm_max', 'm_cnt', 'm_fin_lf', 'm_fin_n'])}


def _push(a: np.ndarray, x: float):
    """`kpi_push` of the CUDA kernel on a copy of the accumulators `[E, NKE]` (used for the baseline's trailing entry)."""
    a = a.copy()
    n = a[:, KE['n']]
    a[:, KE['ramp']] += np.where(n > 0, np.maximum(x - a[:, KE['prev']], 0.0), 0.0)
    a[:, KE['prev']] = x
    a[:, KE['all_max']] = np.where(n > 0, np.maximum(a[:, KE['all_max']], x), x)
    a[:, KE['n']] = n + 1
    for p, w in (('d', 24.0), ('m', 730.0)):
        a[:, KE[f'{p}_sum']] += x
        a[:, KE[f'{p}_max']] = np.where(a[:, KE[f'{p}_cnt']] > 0, np.maximum(a[:, KE[f'{p}_max']], x), x)
        a[:, KE[f'{p}_cnt']] += 1
        full = a[:, KE[f'{p}_cnt']] == w
        with np.errstate(invalid='ignore', divide='ignore'):
            a[:, KE[f'{p}_fin_lf']] += np.where(full, 1.0 - (a[:, KE[f'{p}_sum']] / w) / a[:, KE[f'{p}_max']], 0.0)
        if p == 'd':
            a[:, KE['d_fin_peak']] += np.where(full, a[:, KE['d_max']], 0.0)
        a[:, KE[f'{p}_fin_n']] += full
        a[:, KE[f'{p}_sum']] = np.where(full, 0.0, a[:, KE[f'{p}_sum']])
        a[:, KE[f'{p}_cnt']] = np.where(full, 0.0, a[:, KE[f'{p}_cnt']])
    return a


def _windowed(a: np.ndarray, p: str, what: str) -> np.ndarray:
    """mean over the windows (finished ones + the open one) of `1 - mean/max` ('lf') or of the window maximum ('peak')."""
    cnt = a[:, KE[f'{p}_cnt']]
    open_ = cnt > 0
    with np.errstate(invalid='ignore', divide='ignore'):
        if what == 'lf':
            last = 1.0 - (a[:, KE[f'{p}_sum']] / np.maximum(cnt, 1.0)) / a[:, KE[f'{p}_max']]
            total = a[:, KE[f'{p}_fin_lf']] + np.where(open_, last, 0.0)
        else:
            total = a[:, KE['d_fin_peak']] + np.where(open_, a[:, KE['d_max']], 0.0)
        return total / (a[:, KE[f'{p}_fin_n']] + open_)


def _safe_div_array(c: np.ndarray, b: np.ndarray) -> np.ndarray:
    c = np.where(np.isfinite(c), c, 0.0)
    b = np.where(np.isfinite(b), b, 0.0)
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.where(b == 0.0, np.where(c == 0.0, 1.0, np.nan), c / np.where(b == 0.0, 1.0, b))


def evaluate_batched(spec: S.DistrictSpec, unit: np.ndarray, env: np.ndarray) -> Dict[str, Dict[str, np.ndarray]]:
    """KPI ratios of every env from the accumulators `unit [E, B, 8]`, `env [E, 2, 15]` (control vs `_without_storage` baseline).

    Returns `{'district': {cost_function: [E]}, 'building': {cost_function: [E, B]}}` with the action-dependent rows of
    `CityLearnEnv.evaluate()`: electricity_consumption_total, zero_net_energy, carbon_emissions_total, cost_total (building level,
    district = mean over buildings) and ramping_average, daily / monthly_one_minus_load_factor_average, daily_peak_average,
    all_time_peak_average (district level).  Same slicing quirk as the history path: the baseline district series carries one
    trailing not-yet-simulated (zero) entry, the control series does not (citylearn.py:1188-1200).
    """
    unit = np.asarray(unit, dtype='float64')
    env = np.asarray(env, dtype='float64')
    c, b = env[:, 0, :], _push(env[:, 1, :], 0.0)
    has_carbon = np.array([float(x.series['carbon_intensity'].sum()) != 0 for x in spec.buildings])
    has_price = np.array([float(x.series['electricity_pricing'].sum()) != 0 for x in spec.buildings])
    building = {
        'electricity_consumption_total': _safe_div_array(unit[..., KU['ec']], unit[..., KU['b_ec']]),
        'zero_net_energy': _safe_div_array(unit[..., KU['zne']], unit[..., KU['b_zne']]),
        'carbon_emissions_total': _safe_div_array(unit[..., KU['emission']], np.where(has_carbon[None, :], unit[..., KU['b_emission']], 0.0)),
        'cost_total': _safe_div_array(unit[..., KU['cost']], np.where(has_price[None, :], unit[..., KU['b_cost']], 0.0)),
    }
    district = {k: np.nanmean(v, axis=1) for k, v in building.items()}
    district.update({
        'ramping_average': _safe_div_array(c[:, KE['ramp']], b[:, KE['ramp']]),
        'daily_one_minus_load_factor_average': _safe_div_array(_windowed(c, 'd', 'lf'), _windowed(b, 'd', 'lf')),
        'monthly_one_minus_load_factor_average': _safe_div_array(_windowed(c, 'm', 'lf'), _windowed(b, 'm', 'lf')),
        'daily_peak_average': _safe_div_array(_windowed(c, 'd', 'peak'), _windowed(b, 'd', 'peak')),
        'all_time_peak_average': _safe_div_array(c[:, KE['all_max']], b[:, KE['all_max']]),
    })
    return {'district': district, 'building': building}
