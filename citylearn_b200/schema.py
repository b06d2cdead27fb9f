"""Host-side schema loader: schema.json (+ overrides) -> `DistrictSpec` (flat, device-ready description).

This is the only place that understands the reference's schema format.  It mirrors, function by
function, what the reference does at construction time, but produces struct-of-arrays parameters
and one time-major float32 table instead of a Python object graph:

* override precedence and building inclusion ......... `citylearn/citylearn.py:1973-2086` (`_load`)
* per-building file loading and time-series ingestion . `citylearn/citylearn.py:2172-2207`, `citylearn/data.py:399-661`
* observation / action metadata ....................... `citylearn/citylearn.py:2411-2555` (`process_metadata`)
* md5-derived device seeds and stochastic defaults ..... `citylearn/citylearn.py:2364-2378`, `citylearn/energy_model.py:65-83,194-207,373-376,686-701,960-1012`
* observation / action space estimation ................ `citylearn/building.py:1836-2106,2161-2282`
* episode windows ....................................... `citylearn/base.py:76-129`
* power-outage signals .................................. `citylearn/power_outage.py:27-53,120-169`, `citylearn/building.py:2566-2594`

Nothing in this module touches the GPU.
"""
from __future__ import annotations

import copy
import hashlib
import json
import math
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, List, Mapping, Optional, Tuple, Union

import numpy as np

from .data import DataSet, DataSource, DirectorySource, ZERO_DIVISION_PLACEHOLDER

EPS = ZERO_DIVISION_PLACEHOLDER

# ------------------------------------------------------------------------------------------------
# constants shared with the C ABI (include/citylearn_b200.h).  tests/test_abi.py parses the header
# and checks that these agree.
# ------------------------------------------------------------------------------------------------

# float parameters per building  (cl_building_param)
P = {name: i for i, name in enumerate([
    # battery (citylearn/energy_model.py:872-1242)
    'BAT_CAPACITY', 'BAT_NOMINAL_POWER', 'BAT_EFFICIENCY0', 'BAT_LOSS', 'BAT_CLC', 'BAT_DOD', 'BAT_INITIAL_SOC',
    # storage tanks: cooling, heating, dhw (citylearn/energy_model.py:603-870)
    'CS_CAPACITY', 'CS_EFFICIENCY', 'CS_LOSS', 'CS_INITIAL_SOC', 'CS_MAX_IN', 'CS_MAX_OUT',
    'HS_CAPACITY', 'HS_EFFICIENCY', 'HS_LOSS', 'HS_INITIAL_SOC', 'HS_MAX_IN', 'HS_MAX_OUT',
    'DS_CAPACITY', 'DS_EFFICIENCY', 'DS_LOSS', 'DS_INITIAL_SOC', 'DS_MAX_IN', 'DS_MAX_OUT',
    # cooling device: always a heat pump (citylearn/energy_model.py:157-352)
    'CD_NOMINAL_POWER', 'CD_COP_NUM', 'CD_TARGET',
    # heating / dhw device: heat pump (COP_NUM, TARGET) or electric heater (EFFICIENCY)
    'HD_NOMINAL_POWER', 'HD_COP_NUM', 'HD_TARGET', 'HD_EFFICIENCY',
    'DD_NOMINAL_POWER', 'DD_COP_NUM', 'DD_TARGET', 'DD_EFFICIENCY',
    # time scaling (citylearn/data.py:428-455, citylearn/building.py:113)
    'TIME_STEP_RATIO', 'HOURS_PER_STEP',
    # battery curves: up to 8 points each (x then y)
    'PE_X0', 'PE_X1', 'PE_X2', 'PE_X3', 'PE_X4', 'PE_X5', 'PE_X6', 'PE_X7',
    'PE_Y0', 'PE_Y1', 'PE_Y2', 'PE_Y3', 'PE_Y4', 'PE_Y5', 'PE_Y6', 'PE_Y7',
    'CP_X0', 'CP_X1', 'CP_X2', 'CP_X3', 'CP_X4', 'CP_X5', 'CP_X6', 'CP_X7',
    'CP_Y0', 'CP_Y1', 'CP_Y2', 'CP_Y3', 'CP_Y4', 'CP_Y5', 'CP_Y6', 'CP_Y7',
    # LSTM output de-normalisation for indoor_dry_bulb_temperature (citylearn/building.py:3031-3037)
    'DYN_TIN_MIN', 'DYN_TIN_MAX', 'DYN_CDEM_MIN', 'DYN_CDEM_MAX',
    'PV_NOMINAL_POWER',
])}
NPARAM = len(P)
MAX_CURVE = 8

# int parameters per building  (cl_building_iparam)
IP = {name: i for i, name in enumerate([
    'FLAGS', 'PE_N', 'CP_N',
    # action slot of each action inside the district action vector (-1: inactive)
    'A_COOLING_DEVICE', 'A_HEATING_DEVICE', 'A_COOLING_OR_HEATING_DEVICE',
    'A_COOLING_STORAGE', 'A_HEATING_STORAGE', 'A_DHW_STORAGE', 'A_ELECTRICAL_STORAGE',
    # table columns (index into one time row)
    'C_NSL', 'C_DHW_DEMAND', 'C_COOLING_DEMAND', 'C_HEATING_DEMAND', 'C_SOLAR', 'C_T_OUT', 'C_PRICE', 'C_CARBON',
    'C_HVAC_MODE', 'C_T_IN', 'C_COOL_SP', 'C_HEAT_SP', 'C_COMFORT_BAND', 'C_OCCUPANT',
    # LSTM dynamics: first of the pre-normalised exogenous input columns, slots of the two fed-back inputs, weight offset
    'DYN_C_INPUTS', 'DYN_N_INPUTS', 'DYN_SLOT_TIN', 'DYN_SLOT_CDEM', 'DYN_W_OFFSET', 'DYN_LOOKBACK', 'DYN_HIDDEN',
])}
NIPARAM = len(IP)

# FLAGS bits
F_HEATING_IS_HEAT_PUMP = 1 << 0
F_DHW_IS_HEAT_PUMP = 1 << 1
F_SIMULATE_OUTAGE = 1 << 2
F_DYNAMICS = 1 << 3
F_HAS_THERMAL = 1 << 4          # any thermal device/tank/demand present -> thermal path must run
F_CS_HAS_MAX_IN = 1 << 5
F_CS_HAS_MAX_OUT = 1 << 6
F_HS_HAS_MAX_IN = 1 << 7
F_HS_HAS_MAX_OUT = 1 << 8
F_DS_HAS_MAX_IN = 1 << 9
F_DS_HAS_MAX_OUT = 1 << 10
F_CS_CAPACITY_F32 = 1 << 11     # autosized tank: capacity is an np.float32 in the reference -> `action * capacity` is a float32 product
F_HS_CAPACITY_F32 = 1 << 12
F_CD_NOMINAL_F32 = 1 << 13      # autosized heat pump: np.float32 nominal power -> float32 `action * nominal_power`
F_HD_NOMINAL_F32 = 1 << 14

# per-unit dynamic values (obs writer / reward / trace)  (cl_dyn)
DYN = {name: i for i, name in enumerate([
    'electrical_storage_soc', 'cooling_storage_soc', 'heating_storage_soc', 'dhw_storage_soc',
    'net_electricity_consumption', 'cooling_demand', 'heating_demand', 'dhw_demand',
    'cooling_electricity_consumption', 'heating_electricity_consumption', 'dhw_electricity_consumption',
    'cooling_storage_electricity_consumption', 'heating_storage_electricity_consumption',
    'dhw_storage_electricity_consumption', 'electrical_storage_electricity_consumption',
    'indoor_dry_bulb_temperature', 'non_shiftable_load_electricity_consumption',
    'electrical_storage_energy_balance', 'cooling_storage_energy_balance', 'heating_storage_energy_balance',
    'dhw_storage_energy_balance', 'net_electricity_consumption_cost', 'net_electricity_consumption_emission',
    'electrical_storage_degraded_capacity',
    'energy_to_non_shiftable_load', 'cooling_demand_series', 'heating_demand_series',
])}
NDYN = len(DYN)

# observation descriptor kinds  (cl_obs_kind)
OBS_TS, OBS_DYN, OBS_OUTAGE, OBS_TS_MINUS_TS = 0, 1, 2, 3
OBS_STATE = 4            # a = slot of the per-env charging-constraint state (cl_ev_desc): headroom / violation of the last applied actions
CC_SLOTS = 6             # per constrained building: building headroom, 4 phase headrooms, violation (kWh)

# built-in reward ids  (cl_reward_id)
REWARD_IDS = {
    'RewardFunction': 0, 'MARL': 1, 'IndependentSACReward': 2, 'SolarPenaltyReward': 3,
    'ComfortReward': 4, 'SolarPenaltyAndComfortReward': 5, 'custom': -1,
}

PERIODIC = {'hour': 24, 'day_type': 7, 'month': 12, 'minutes': 60}  # citylearn/building.py:1483-1498

ENERGY_SIMULATION_COLUMNS = [
    'month', 'hour', 'day_type', 'daylight_savings_status', 'indoor_dry_bulb_temperature',
    'average_unmet_cooling_setpoint_difference', 'indoor_relative_humidity', 'non_shiftable_load', 'dhw_demand',
    'cooling_demand', 'heating_demand', 'solar_generation', 'occupant_count',
    'indoor_dry_bulb_temperature_cooling_set_point', 'indoor_dry_bulb_temperature_heating_set_point', 'hvac_mode',
    'power_outage', 'comfort_band', 'minutes',
]
INT_COLUMNS = {'month', 'hour', 'day_type', 'daylight_savings_status', 'hvac_mode', 'minutes'}
WEATHER_COLUMNS = [
    'outdoor_dry_bulb_temperature', 'outdoor_relative_humidity', 'diffuse_solar_irradiance', 'direct_solar_irradiance',
    'outdoor_dry_bulb_temperature_predicted_1', 'outdoor_dry_bulb_temperature_predicted_2', 'outdoor_dry_bulb_temperature_predicted_3',
    'outdoor_relative_humidity_predicted_1', 'outdoor_relative_humidity_predicted_2', 'outdoor_relative_humidity_predicted_3',
    'diffuse_solar_irradiance_predicted_1', 'diffuse_solar_irradiance_predicted_2', 'diffuse_solar_irradiance_predicted_3',
    'direct_solar_irradiance_predicted_1', 'direct_solar_irradiance_predicted_2', 'direct_solar_irradiance_predicted_3',
]
PRICING_COLUMNS = ['electricity_pricing', 'electricity_pricing_predicted_1', 'electricity_pricing_predicted_2', 'electricity_pricing_predicted_3']
CARBON_COLUMNS = ['carbon_intensity']

SUPPORTED_BUILDING_TYPES = {
    'citylearn.citylearn.Building': False, 'citylearn.building.Building': False,
    'citylearn.building.DynamicsBuilding': True, 'citylearn.building.LSTMDynamicsBuilding': True,
}


class UnknownSchemaError(Exception):
    """Mirrors `citylearn.citylearn.UnknownSchemaError` (`citylearn/citylearn.py:2679`)."""

    def __init__(self, message=None):
        super().__init__('Unknown schema parsed into constructor. Schema must be name of CityLearn data set,'
                         ' a filepath to JSON representation or `dict` object of a CityLearn schema.' if message is None else message)


class UnsupportedSchemaError(NotImplementedError):
    """The schema asks for a reference feature outside the accelerated hot path (EVs, washing machines, occupants...)."""


# ------------------------------------------------------------------------------------------------
# episode tracker (citylearn/base.py:6-134)
# ------------------------------------------------------------------------------------------------
class EpisodeTracker:
    def __init__(self, simulation_start_time_step: int, simulation_end_time_step: int):
        self.simulation_start_time_step = int(simulation_start_time_step)
        self.simulation_end_time_step = int(simulation_end_time_step)
        self.episode_start_time_step = None
        self.episode_end_time_step = None
        self.reset_episode_index()

    @property
    def episode_time_steps(self) -> int:
        return self.episode_end_time_step - self.episode_start_time_step + 1

    @property
    def simulation_time_steps(self) -> int:
        return self.simulation_end_time_step - self.simulation_start_time_step + 1

    def splits(self, episode_time_steps, rolling_episode_split: bool) -> List[Tuple[int, int]]:
        if isinstance(episode_time_steps, list):
            return [tuple(s) for s in episode_time_steps]
        earliest = self.simulation_start_time_step
        latest = self.simulation_end_time_step + 1 - episode_time_steps
        step = 1 if rolling_episode_split else episode_time_steps
        starts = range(earliest, latest + 1, step)
        return [(s, s + episode_time_steps - 1) for s in starts]

    def pick(self, episode: int, episode_time_steps, rolling_episode_split: bool, random_episode_split: bool, random_seed: int) -> Tuple[int, int]:
        """Window of episode number `episode` (citylearn/base.py:84-126)."""
        splits = self.splits(episode_time_steps, rolling_episode_split)
        if random_episode_split:
            seed = int(random_seed * (episode + 1))
            ix = np.random.RandomState(seed).choice(len(splits) - 1)
        else:
            ix = episode % len(splits)
        return splits[ix]

    def next_episode(self, episode_time_steps, rolling_episode_split: bool, random_episode_split: bool, random_seed: int):
        self.episode += 1
        self.episode_start_time_step, self.episode_end_time_step = self.pick(
            self.episode, episode_time_steps, rolling_episode_split, random_episode_split, random_seed)

    def reset_episode_index(self):
        self.episode = -1


# ------------------------------------------------------------------------------------------------
# power outage signal generators (citylearn/power_outage.py)
# ------------------------------------------------------------------------------------------------
class PowerOutage:
    def __init__(self, random_seed: int = None):
        self._random_seed = random_seed

    @property
    def random_seed(self) -> int:
        return np.random.randint(0, 100_000_000) if self._random_seed is None else self._random_seed

    def get_signals(self, time_steps: int, **kwargs) -> np.ndarray:
        return np.random.RandomState(self.random_seed).choice([0, 1], size=time_steps)


class ReliabilityMetricsPowerOutage(PowerOutage):
    def __init__(self, saifi: float = None, caidi: float = None, start_time_steps: List[int] = None, **kwargs):
        super().__init__(**kwargs)
        self.saifi = 1.436 if saifi is None else saifi
        self.caidi = 331.2 if caidi is None else caidi
        self.start_time_steps = start_time_steps

    def get_signals(self, time_steps: int, seconds_per_time_step: float = 3600.0, **kwargs) -> np.ndarray:
        nprs = np.random.RandomState(self.random_seed)
        time_steps_per_day = 86400.0 / seconds_per_time_step
        time_steps_per_minute = 60.0 / seconds_per_time_step
        day_count = time_steps / time_steps_per_day
        outage_days = nprs.binomial(n=1, p=self.saifi / 365.0, size=int(day_count))
        outage_day_ixs = outage_days * np.arange(day_count)
        outage_day_ixs = outage_day_ixs[outage_day_ixs != 0]
        outage_day_count = outage_days[outage_days == 1].shape[0]
        starts = list(range(int(time_steps_per_day))) if self.start_time_steps is None else self.start_time_steps
        outage_starts = nprs.choice(starts, size=outage_day_count)
        durations = nprs.exponential(scale=self.caidi, size=outage_day_count) * time_steps_per_minute
        signals = np.zeros(time_steps, dtype=int)
        for i, j, k in zip(outage_day_ixs, outage_starts, durations):
            start_ix = i * time_steps_per_day + j
            signals[int(start_ix):int(start_ix + k)] = 1
        return signals


_OUTAGE_MODELS = {
    'citylearn.power_outage.PowerOutage': PowerOutage,
    'citylearn.power_outage.ReliabilityMetricsPowerOutage': ReliabilityMetricsPowerOutage,
    'citylearn_b200.schema.PowerOutage': PowerOutage,
    'citylearn_b200.schema.ReliabilityMetricsPowerOutage': ReliabilityMetricsPowerOutage,
}


# ------------------------------------------------------------------------------------------------
# spec dataclasses
# ------------------------------------------------------------------------------------------------
@dataclass
class BuildingSpec:
    name: str
    index: int
    building_type: str
    dynamics: bool
    observation_metadata: Dict[str, bool]
    action_metadata: Dict[str, bool]
    series: Dict[str, np.ndarray]                 # full-length ingested series (float32 / int32), reference dtypes
    devices: Dict[str, Dict[str, Any]]            # resolved device attributes (after stochastic defaults)
    time_step_ratio: float
    seconds_per_time_step: float
    simulate_power_outage: bool
    stochastic_power_outage: bool
    outage_model: Optional[PowerOutage]
    dynamics_attrs: Optional[Dict[str, Any]] = None
    dynamics_weights: Optional[Dict[str, np.ndarray]] = None
    observation_low: Optional[np.ndarray] = None
    observation_high: Optional[np.ndarray] = None
    action_low: Optional[np.ndarray] = None
    action_high: Optional[np.ndarray] = None
    maximum_temperature_delta: float = 20.0
    observation_space_limit_delta: float = 0.0
    demand_observation_limit_factor: float = 2.0
    chargers: List[Any] = field(default_factory=list)            # ev.ChargerSpec
    washing_machines: List[Any] = field(default_factory=list)    # ev.WashingMachineSpec
    charging_constraints: Optional[Any] = None                   # ev.ChargingConstraints (citylearn/building.py:764-989)
    # names of the observation VALUES in the order `Building.observations()` returns them, when it differs from `active_observations`
    # (the names / space order): charging-constraint observations precede the per-charger ones in the values (building.py:1146-1154)
    observation_value_order: Optional[List[str]] = None

    @property
    def active_observations(self) -> List[str]:
        return [k for k, v in self.observation_metadata.items() if v]

    @property
    def active_actions(self) -> List[str]:
        return [k for k, v in self.action_metadata.items() if v]


@dataclass
class DistrictSpec:
    buildings: List[BuildingSpec]
    schema: dict
    root_directory: Optional[str]
    central_agent: bool
    shared_observations: List[str]
    random_seed: int
    seconds_per_time_step: float
    episode_time_steps: Any
    rolling_episode_split: bool
    random_episode_split: bool
    simulation_start_time_step: int
    simulation_end_time_step: int
    reward_type: Any
    reward_attributes: Any
    time_step_ratio: float = 1.0
    # device-ready arrays, filled by `finalize`
    table: Optional[np.ndarray] = None            # [N, W] float32, one row per dataset time step
    columns: Dict[Any, int] = field(default_factory=dict)
    params: Optional[np.ndarray] = None           # [B, NPARAM] float32
    iparams: Optional[np.ndarray] = None          # [B, NIPARAM] int32
    lstm_weights: Optional[np.ndarray] = None     # flat float32, per-building blocks
    action_dim: int = 0
    evs: List[Any] = field(default_factory=list)   # ev.ElectricVehicleSpec: vehicles shared by the district's chargers
    ev: Optional[Dict[str, Any]] = None           # device-ready electric-vehicle / charger / washing-machine arrays (finalize)
    ev_random_seed: Optional[int] = None

    @property
    def n_buildings(self) -> int:
        return len(self.buildings)


# ------------------------------------------------------------------------------------------------
# helpers reproducing reference device construction
# ------------------------------------------------------------------------------------------------
def device_random_seed(building_name: str, building_type: str, device_name: str, device_type: str, schema_seed: int) -> int:
    """citylearn/citylearn.py:2364-2373: ONE running md5 object, digests of the cumulative prefixes are summed."""
    md5 = hashlib.md5()
    s = 0
    for string in (building_name, building_type, device_name, device_type):
        md5.update(string.encode())
        s += int(md5.hexdigest(), 16)
    return int(str(s * (schema_seed + 1))[:9])


def _draw(value, default, seed):
    """`Device._get_property_value` (citylearn/energy_model.py:65-83): every draw re-seeds RandomState(seed)."""
    if value is None or (isinstance(value, float) and math.isnan(value)):
        if isinstance(default, tuple):
            return float(np.random.RandomState(seed).uniform(*default))
        return default
    if isinstance(value, (tuple, list)) and len(value) == 2 and not isinstance(value[0], (list, tuple)):
        return float(np.random.RandomState(seed).uniform(*value))
    return value


def _absent_seed(name: str, device: str) -> int:
    # The reference seeds absent-device defaults from Python's global `random` (citylearn/base.py:189-191), i.e. they are
    # not reproducible there.  They only parameterise zero-sized devices; use a stable hash so runs are repeatable.
    return int(hashlib.md5(f'{name}/{device}'.encode()).hexdigest()[:7], 16)


def resolve_heat_pump(attrs: Optional[dict], seed: int) -> dict:
    a = dict(attrs or {})
    return {
        'type': 'HeatPump',
        'nominal_power': 0.0 if a.get('nominal_power') is None else float(a['nominal_power']),
        'efficiency': _draw(a.get('efficiency'), (0.2, 0.3), seed),
        'target_heating_temperature': _draw(a.get('target_heating_temperature'), (45.0, 50.0), seed),
        'target_cooling_temperature': _draw(a.get('target_cooling_temperature'), (7.0, 10.0), seed),
    }


def resolve_electric_heater(attrs: Optional[dict], seed: int) -> dict:
    a = dict(attrs or {})
    return {
        'type': 'ElectricHeater',
        'nominal_power': 0.0 if a.get('nominal_power') is None else float(a['nominal_power']),
        'efficiency': _draw(a.get('efficiency'), (0.9, 0.99), seed),
    }


def resolve_storage_tank(attrs: Optional[dict], seed: int) -> dict:
    a = dict(attrs or {})
    return {
        'type': 'StorageTank',
        'capacity': 0.0 if a.get('capacity') is None else float(a['capacity']),
        'efficiency': _draw(a.get('efficiency'), (0.9, 0.98), seed),
        'loss_coefficient': _draw(a.get('loss_coefficient'), (0.001, 0.009), seed),
        'initial_soc': _draw(a.get('initial_soc'), 0.0, seed),
        'max_input_power': a.get('max_input_power'),
        'max_output_power': a.get('max_output_power'),
    }


def resolve_battery(attrs: Optional[dict], seed: int) -> dict:
    a = dict(attrs or {})
    dod = _draw(a.get('depth_of_discharge'), 1.0, seed)
    eff = _draw(a.get('efficiency'), (0.9, 0.98), seed)
    initial_soc = a.get('initial_soc')
    initial_soc = 1.0 - dod if initial_soc is None else _draw(initial_soc, 0.0, seed)
    u = lambda lo, hi: float(np.random.RandomState(seed).uniform(lo, hi))  # noqa: E731
    pe = a.get('power_efficiency_curve')
    if pe is None:  # citylearn/energy_model.py:977-990
        pe = [[0, u(eff * 0.85, eff * 0.9)], [u(0.25, 0.35), u(eff * 0.9, eff * 0.95)], [u(0.65, 0.75), u(eff * 0.98, eff * 1.0)],
              [u(0.75, 0.85), eff], [1, u(eff * 0.95, eff * 0.98)]]
    cp = a.get('capacity_power_curve')
    if cp is None:  # citylearn/energy_model.py:992-1003
        cp = [[0.0, u(0.95, 1.0)], [u(0.75, 0.85), u(0.9, 0.95)], [1.0, u(0.2, 0.3)]]
    pe = np.array(pe, dtype='float64').T
    cp = np.array(cp, dtype='float64').T
    if pe.shape[1] > MAX_CURVE or cp.shape[1] > MAX_CURVE:
        raise UnsupportedSchemaError(f'battery curves with more than {MAX_CURVE} points are not supported')
    return {
        'type': 'Battery',
        'capacity': 0.0 if a.get('capacity') is None else float(a['capacity']),
        'nominal_power': 0.0 if a.get('nominal_power') is None else float(a['nominal_power']),
        'efficiency': eff,
        'loss_coefficient': _draw(a.get('loss_coefficient'), (0.001, 0.009), seed),
        'capacity_loss_coefficient': _draw(a.get('capacity_loss_coefficient'), (1e-05, 0.0001), seed),
        'depth_of_discharge': dod,
        'initial_soc': initial_soc,
        'power_efficiency_curve': pe,
        'capacity_power_curve': cp,
    }


def _autosize(device_name: str, dev: dict, series: Mapping[str, np.ndarray], spec: 'DistrictSpec', seed: int, kwargs: dict):
    """Load-time sizing of heat pumps, heaters and tanks to the building's peak demand over the simulation window
    (`Building.autosize_*`, citylearn/building.py:2284-2403; device formulas citylearn/energy_model.py:309-352, 425-450, 770-795).

    Devices are sized right after construction, while their own `time_step_ratio` is still 1, on float32 series - the sized
    value is therefore a float32 number, which is kept (it feeds float64 arithmetic later).
    """
    window = slice(spec.simulation_start_time_step, spec.simulation_end_time_step + 1)
    end_use = device_name.split('_')[0]
    demand = np.asarray(series[f'{end_use}_demand'][window], dtype='float32')
    with np.errstate(divide='ignore', invalid='ignore'):
        if dev['type'] == 'StorageTank':
            safety = _draw(kwargs.get('safety_factor'), (1.0, 2.0), seed)
            dev['capacity'] = float(np.nanmax(demand * 1.0) * safety)
        elif dev['type'] == 'HeatPump':
            safety = _draw(kwargs.get('safety_factor'), 1.0, seed)
            cop = cop32(dev, series['outdoor_dry_bulb_temperature'][window], heating=end_use != 'cooling')
            dev['nominal_power'] = float(np.nanmax(np.array(demand * 1) / cop + 0) * safety)
        else:
            safety = _draw(kwargs.get('safety_factor'), 1.0, seed)
            dev['nominal_power'] = float(np.nanmax(np.array(demand * 1) / dev['efficiency']) * safety)
    dev['autosized'] = True


_DEVICE_RESOLVERS = {
    'citylearn.energy_model.HeatPump': resolve_heat_pump,
    'citylearn.energy_model.ElectricHeater': resolve_electric_heater,
    'citylearn.energy_model.StorageTank': resolve_storage_tank,
    'citylearn.energy_model.Battery': resolve_battery,
}
_ABSENT_DEFAULT = {  # citylearn/building.py:717-747
    'cooling_device': resolve_heat_pump, 'heating_device': resolve_heat_pump, 'dhw_device': resolve_electric_heater,
    'cooling_storage': resolve_storage_tank, 'heating_storage': resolve_storage_tank, 'dhw_storage': resolve_storage_tank,
    'electrical_storage': resolve_battery,
}


def cop32(dev: dict, temperature: np.ndarray, heating: bool) -> np.ndarray:
    """`HeatPump.get_cop` on a float32 series (citylearn/energy_model.py:216-250), NumPy-2 weak-scalar arithmetic."""
    t = np.asarray(temperature, dtype='float32')
    target = dev['target_heating_temperature'] if heating else dev['target_cooling_temperature']
    num = dev['efficiency'] * (target + 273.15)
    with np.errstate(divide='ignore', invalid='ignore'):
        cop = num / ((target - t) if heating else (t - target))
    cop = np.array(cop)
    cop[cop < 0] = 20
    cop[cop > 20] = 20
    return cop


# ------------------------------------------------------------------------------------------------
# loader
# ------------------------------------------------------------------------------------------------
def resolve_source(schema: Union[str, os.PathLike, Mapping[str, Any]], root_directory=None,
                   data_source: Optional[DataSource] = None) -> Tuple[dict, DataSource]:
    """`CityLearnEnv.schema.setter` (citylearn/citylearn.py:862-883) without the network.

    `data_source` lets a schema *dict* (e.g. a bundled schema with an edited reward_function) read its files from a pack.
    """
    if data_source is not None and isinstance(schema, dict):
        return copy.deepcopy(schema), data_source
    if isinstance(schema, (str, Path)) and os.path.isfile(schema):
        path = Path(schema)
        with open(path) as f:
            sch = json.load(f)
        if sch.get('root_directory') is None:
            sch['root_directory'] = str(path.parent.absolute())
        root = root_directory if root_directory is not None else sch['root_directory']
        return sch, DirectorySource(root, sch)
    if isinstance(schema, str):
        if schema in DataSet.get_dataset_names():
            src = DataSet.get_source(schema)
            sch = src.schema()
            if root_directory is not None:
                return sch, DirectorySource(root_directory, sch)
            return sch, src
        raise UnknownSchemaError()
    if isinstance(schema, dict):
        sch = copy.deepcopy(schema)
        root = root_directory if root_directory is not None else sch.get('root_directory')
        if root is None:
            raise UnknownSchemaError('schema dict needs a root_directory (or pass root_directory=...).')
        return sch, DirectorySource(root, sch)
    raise UnknownSchemaError()


def _ingest_energy_simulation(tab: Dict[str, np.ndarray], seconds_per_time_step: float) -> Tuple[Dict[str, np.ndarray], float]:
    """`EnergySimulation.__init__` (citylearn/data.py:399-493), noise_std = 0."""
    n = len(tab['solar_generation'])
    s: Dict[str, np.ndarray] = {}
    for c in ('month', 'hour', 'day_type'):
        s[c] = np.array(tab[c], dtype='int32')
    s['indoor_dry_bulb_temperature'] = np.clip(np.array(tab['indoor_dry_bulb_temperature'], dtype='float32'), -90, 57)
    for c in ('non_shiftable_load', 'dhw_demand', 'cooling_demand', 'heating_demand', 'solar_generation'):
        s[c] = np.array(tab[c], dtype='float32')
    assert (s['cooling_demand'] * s['heating_demand']).sum() == 0, 'Cooling and heating in the same time step is not allowed.'
    minutes = tab.get('minutes')
    s['minutes'] = None if minutes is None else np.array(minutes, dtype='int32')
    time_delta = int(s['hour'][1]) * 60 - int(s['hour'][0]) * 60
    if s['minutes'] is not None and len(s['minutes']) > 1:
        time_delta = (int(s['hour'][1]) * 60 + int(s['minutes'][1])) - (int(s['hour'][0]) * 60 + int(s['minutes'][0]))
    if time_delta < 0:
        time_delta += 1440
    base_step_seconds = max(1, time_delta * 60)
    ratio = seconds_per_time_step / base_step_seconds if seconds_per_time_step and base_step_seconds else None

    def opt(name, default, dtype='float32', clip=None):
        v = tab.get(name)
        if v is None:
            a = np.zeros(n, dtype=dtype) + default
            return a.astype(dtype)
        a = np.array(v, dtype=dtype)
        return np.clip(a, *clip) if clip is not None else a

    s['daylight_savings_status'] = opt('daylight_savings_status', 0, 'int32')
    s['average_unmet_cooling_setpoint_difference'] = opt('average_unmet_cooling_setpoint_difference', 0.0)
    s['indoor_relative_humidity'] = opt('indoor_relative_humidity', 0.0, clip=(0, 100))
    s['occupant_count'] = opt('occupant_count', 0.0)
    s['indoor_dry_bulb_temperature_cooling_set_point'] = opt('indoor_dry_bulb_temperature_cooling_set_point', 0.0)
    s['indoor_dry_bulb_temperature_heating_set_point'] = opt('indoor_dry_bulb_temperature_heating_set_point', 0.0)
    s['power_outage'] = opt('power_outage', 0.0)
    s['comfort_band'] = opt('comfort_band', 2.0)
    hv = tab.get('hvac_mode')
    if hv is None:
        s['hvac_mode'] = np.zeros(n, dtype='int32') + 1
    else:
        bad = set(np.unique(hv[~np.isnan(hv)]).tolist()) - {0, 1, 2, 3}
        assert not bad, f'Invalid hvac_mode values were found: {sorted(bad)}.'
        s['hvac_mode'] = np.array(hv, dtype='int32')
    return s, ratio


def load(schema: Union[str, os.PathLike, Mapping[str, Any]], **kwargs) -> DistrictSpec:
    """schema (+ `CityLearnEnv.__init__` keyword overrides) -> finalized `DistrictSpec`.

    An already loaded `DistrictSpec` is returned as is (several envs over one large district); overrides need a reload."""
    if isinstance(schema, DistrictSpec):
        extra = [k for k, v in kwargs.items() if v is not None]
        if extra:
            raise ValueError(f'a loaded DistrictSpec cannot take overrides {extra}: load the schema again with them')
        return schema
    sch, source = resolve_source(schema, kwargs.get('root_directory'), kwargs.get('data_source'))
    sch = copy.deepcopy(sch)
    g = lambda k: kwargs.get(k)  # noqa: E731

    # ---- citylearn/citylearn.py:2006-2051 (override precedence) ----
    random_seed = sch.get('random_seed', None)   # NOTE: the `random_seed` kwarg never reaches schema['random_seed'] (:2008)
    env_random_seed = random_seed if g('random_seed') is None else g('random_seed')
    central_agent = g('central_agent') if g('central_agent') is not None else sch['central_agent']
    ev_obs = [k for k in sch['observations'] if 'electric_vehicle_' in k]
    wm_obs = [k for k in sch['observations'] if 'washing_machine_' in k]
    ev_act = [k for k in sch['actions'] if 'electric_vehicle_' in k]
    wm_act = [k for k in sch['actions'] if 'washing_machine' in k]
    observations = {k: v for k, v in sch['observations'].items() if k not in set(ev_obs) | set(wm_obs)}
    actions = {k: v for k, v in sch['actions'].items() if k not in set(ev_act) | set(wm_act)}
    shared_observations = g('shared_observations') if g('shared_observations') is not None else [
        k for k, v in observations.items() if v.get('shared_in_central_agent', False)]
    episode_time_steps = g('episode_time_steps') if g('episode_time_steps') is not None else sch.get('episode_time_steps', None)
    rolling = g('rolling_episode_split') if g('rolling_episode_split') is not None else sch.get('rolling_episode_split', None)
    # the reference forwards this kwarg under the wrong name (`random_episode=`, citylearn/citylearn.py:207), so the
    # schema value always wins there; we honour the kwarg when given (documented deviation).
    random_split = g('random_episode_split') if g('random_episode_split') is not None else sch.get('random_episode_split', None)
    seconds_per_time_step = g('seconds_per_time_step') if g('seconds_per_time_step') is not None else sch['seconds_per_time_step']
    sim_start = g('simulation_start_time_step') if g('simulation_start_time_step') is not None else sch['simulation_start_time_step']
    sim_end = g('simulation_end_time_step') if g('simulation_end_time_step') is not None else sch['simulation_end_time_step']

    # ---- building inclusion (citylearn/citylearn.py:2059-2086) ----
    names = list(sch['buildings'].keys())
    sel = g('buildings')
    if sel is not None and len(sel) > 0:
        if isinstance(sel[0], str):
            names = [b for b in names if b in sel]
        elif isinstance(sel[0], (int, np.integer)):
            names = [names[i] for i in sel]
        else:
            raise Exception('Unknown buildings type. Allowed types are int and str.')
    else:
        names = [b for b in names if sch['buildings'][b]['include']]

    spec = DistrictSpec(
        buildings=[], schema=sch, root_directory=source.root_directory, central_agent=bool(central_agent),
        shared_observations=list(shared_observations), random_seed=env_random_seed,
        seconds_per_time_step=float(seconds_per_time_step),
        episode_time_steps=episode_time_steps, rolling_episode_split=bool(rolling), random_episode_split=bool(random_split),
        simulation_start_time_step=int(sim_start), simulation_end_time_step=int(sim_end),
        reward_type=None, reward_attributes=None)

    # electric vehicles, chargers, washing machines (citylearn/citylearn.py:2010-2016, 2088-2098): per-charger / per-machine
    # observations and actions are expanded from these schema-level switches in `_load_building`
    from . import ev as EV
    spec.evs = EV.load_electric_vehicles(sch, kwargs, resolve_battery)
    spec.ev_random_seed = g('ev_random_seed') if g('ev_random_seed') is not None else random_seed
    spec._helpers = {'ev_obs': {k: sch['observations'][k] for k in ev_obs}, 'wm_obs': {k: sch['observations'][k] for k in wm_obs},
                     'ev_act': {k: sch['actions'][k] for k in ev_act}, 'wm_act': {k: sch['actions'][k] for k in wm_act}}
    ratios: List[float] = []
    for index, name in enumerate(names):
        spec.buildings.append(_load_building(index, name, sch, source, observations, actions, random_seed,
                                             float(seconds_per_time_step), ratios, spec, kwargs))
    spec.time_step_ratio = spec.buildings[0].time_step_ratio if spec.buildings else 1.0

    # ---- reward function selection (citylearn/citylearn.py:2100-2163) ----
    reward_schema = sch['reward_function']
    reward_type = reward_schema['type']
    reward_attrs = reward_schema.get('attributes', {})
    if isinstance(reward_type, dict):
        spec.reward_type, spec.reward_attributes = reward_type, reward_attrs
    else:
        if g('reward_function') is not None:
            rt = g('reward_function')
            if not isinstance(rt, str):
                rt = rt if isinstance(rt, type) else type(rt)
            reward_type = rt
        # quirk kept: schema attributes leak to an overriding class unless reward_function_kwargs is truthy (:2154)
        spec.reward_type = reward_type
        spec.reward_attributes = g('reward_function_kwargs') or reward_attrs or {}

    # vehicle schedule: who is plugged in where, arrival SOCs, away-drift factors - all action-independent (ev.compile_schedule);
    # charger observations become plain series of their building
    all_chargers = [c for b in spec.buildings for c in b.chargers]
    if all_chargers:
        n_rows = len(spec.buildings[0].series['hour'])
        sched = EV.compile_schedule(all_chargers, len(spec.evs), n_rows, spec.simulation_start_time_step, spec.simulation_end_time_step,
                                    spec.ev_random_seed)
        spec.ev = {'schedule': sched}
        init = np.array([e.battery['initial_soc'] for e in spec.evs], dtype='float64')
        for b in spec.buildings:
            for c in b.chargers:
                cols = EV.charger_observation_columns(c, sched, init)
                for key, pattern in EV.CHARGER_OBSERVATIONS:
                    b.series[pattern.format(id=c.charger_id)] = cols[key]
                # the same observation on the FIRST row of an episode: every connection is new at t = 0 (arrival SOC, else the vehicle's
                # initial SOC stays in soc[0])
                con = EV.connected_mask(c)
                t0v = sched['t0'][np.arange(len(con)), np.maximum(c.ev, 0)]
                soc0 = np.where(con, np.where(np.isnan(t0v), init[np.maximum(c.ev, 0)], t0v), EV.DEFAULT_SOC)
                b.series[f'connected_electric_vehicle_at_charger_{c.charger_id}_soc__t0'] = np.asarray(soc0, dtype='float32')
    finalize(spec)
    return spec


def _load_building(index, name, sch, source: DataSource, observations, actions, schema_seed, seconds_per_time_step,
                   ratios, spec: DistrictSpec, kwargs) -> BuildingSpec:
    bs = sch['buildings'][name]
    for unsupported in ('occupant',):
        if bs.get(unsupported):
            raise UnsupportedSchemaError(f"building '{name}': '{unsupported}' is outside the accelerated hot path (SURVEY.md §8f)")
    if bs.get('noise_std', 0.0):
        raise UnsupportedSchemaError('noise_std > 0 is not supported')
    building_type = 'citylearn.citylearn.Building' if bs.get('type') is None else bs['type']
    if building_type not in SUPPORTED_BUILDING_TYPES:
        raise UnsupportedSchemaError(f"building type '{building_type}' is outside the accelerated hot path")
    series, ratio = _ingest_energy_simulation(source.table(bs['energy_simulation']), seconds_per_time_step)
    ratios.append(ratio)
    # quirk kept: `time_step_ratios` is one shared, ever-growing list indexed by building position (citylearn/data.py:403,454)
    time_step_ratio = ratios[index]
    n = len(series['hour'])
    wt = source.table(bs['weather'])
    for c in WEATHER_COLUMNS:
        series[c] = np.array(wt[c], dtype='float32')
    if bs.get('carbon_intensity') is not None:
        series['carbon_intensity'] = np.clip(np.array(source.table(bs['carbon_intensity'])['carbon_intensity'], dtype='float32'), 0, 1)
    else:
        series['carbon_intensity'] = np.zeros(n, dtype='float32')
    if bs.get('pricing') is not None:
        pt = source.table(bs['pricing'])
        for c in PRICING_COLUMNS:
            series[c] = np.clip(np.array(pt[c], dtype='float32'), 0, 1)
    else:
        for c in PRICING_COLUMNS:
            series[c] = np.zeros(n, dtype='float32')

    # ---- metadata (citylearn/citylearn.py:2411-2555) ----
    om = {k: v['active'] for k, v in observations.items()}
    if 'minutes' in om and series['minutes'] is None:
        om.pop('minutes', None)
    ao = kwargs.get('active_observations')
    if ao is not None:
        ao = ao[index] if isinstance(ao[0], list) else ao
        om = {k: k in ao for k in om}
    io_ = kwargs.get('inactive_observations')
    if io_ is not None:
        io_ = io_[index] if isinstance(io_[0], list) else io_
    elif bs.get('inactive_observations') is not None:
        io_ = bs['inactive_observations']
    else:
        io_ = []
    om = {k: False if k in io_ else om[k] for k in om}
    am = {k: v['active'] for k, v in actions.items()}
    aa = kwargs.get('active_actions')
    if aa is not None:
        aa = aa[index] if isinstance(aa[0], list) else aa
        am = {k: k in aa for k in am}
    ia = kwargs.get('inactive_actions')
    if ia is not None:
        ia = ia[index] if isinstance(ia[0], list) else ia
    elif bs.get('inactive_actions') is not None:
        ia = bs['inactive_actions']
    else:
        ia = []
    am = {k: False if k in ia else v for k, v in am.items()}
    # per-charger / per-machine observations and actions (process_metadata, citylearn/citylearn.py:2411-2555)
    from . import ev as EV
    hp = getattr(spec, '_helpers', {'ev_obs': {}, 'wm_obs': {}, 'ev_act': {}, 'wm_act': {}})
    lo_row, hi_row = spec.simulation_start_time_step, spec.simulation_end_time_step
    chargers = EV.load_chargers(index, bs, source, lo_row, hi_row, [e.name for e in spec.evs])
    wms = EV.load_washing_machines(index, bs, source, kwargs)
    for obj, names_ in [(c, ('state', 'ev', 'capacity', 'current_soc', 'departure_time', 'required_soc', 'arrival_time', 'soc_arrival')) for c in chargers] + \
                       [(w, ('start', 'end', 'profile_sum', 'profile_len', 'profile_prefix')) for w in wms]:
        for an in names_:          # schedules index by dataset row like every other series
            a = getattr(obj, an)
            if len(a) < n:
                raise UnsupportedSchemaError(f"building '{name}': a charger / washing-machine schedule is shorter than the building's series")
            setattr(obj, an, a[:n])

    def helper_flags(h, active_list, inactive_list):
        f = {k: v['active'] for k, v in h.items()}
        if active_list is not None:
            f = {k: k in active_list for k in f}
        return {k: False if k in inactive_list else v for k, v in f.items()}
    ev_obs_f, wm_obs_f = helper_flags(hp['ev_obs'], ao, io_), helper_flags(hp['wm_obs'], ao, io_)
    ev_act_f, wm_act_f = helper_flags(hp['ev_act'], aa, ia), helper_flags(hp['wm_act'], aa, ia)
    for c in chargers:
        for key, pattern in EV.CHARGER_OBSERVATIONS:
            if ev_obs_f.get(key, False):
                om[pattern.format(id=c.charger_id)] = True
        if ev_act_f.get('electric_vehicle_storage', False):
            am[f'electric_vehicle_storage_{c.charger_id}'] = True
    for w in wms:
        if wm_obs_f.get('washing_machine_start_time_step', False):
            om[f'{w.name}_start_time_step'] = True
        if wm_obs_f.get('washing_machine_end_time_step', False):
            om[f'{w.name}_end_time_step'] = True
        if wm_act_f.get('washing_machine', False):
            am[f'{w.name}'] = True

    # charging constraints (citylearn/citylearn.py:2177-2178, building.py:764-833): observation names in the reference's insertion order
    cc = EV.load_charging_constraints(bs.get('charging_constraints'), chargers)
    value_order = None
    if cc is not None:
        if len(cc.phases) > EV.MAX_PHASES:
            raise UnsupportedSchemaError(f"building '{name}': more than {EV.MAX_PHASES} charging phases are not supported")
        for k in cc.one_hot:
            om.setdefault(k, True)
        if cc.expose_headroom:
            for k in cc.headroom_names:
                om.setdefault(k, True)
        om['charging_constraint_violation_kwh'] = cc.expose_violation
        for k in cc.one_hot:
            om[k] = True
        # `observations = {k: data[k] for k in valid_observations if k in data}` then the chargers' and machines' values are appended
        per_device = {pattern.format(id=c.charger_id) for c in chargers for _, pattern in EV.CHARGER_OBSERVATIONS} | \
            {f'{w.name}_{k}' for w in wms for k in ('start_time_step', 'end_time_step')}
        active = [k for k, v in om.items() if v]
        value_order = [k for k in active if k not in per_device] + [k for k in active if k in per_device]
        if value_order == active:
            value_order = None

    # ---- power outage (citylearn/citylearn.py:2273-2290) ----
    po = bs.get('power_outage', {}) or {}
    simulate = kwargs.get('simulate_power_outage')
    simulate = po.get('simulate_power_outage') if simulate is None else simulate
    simulate = simulate[index] if isinstance(simulate, list) else simulate
    stochastic = po.get('stochastic_power_outage')
    model = None
    if po.get('stochastic_power_outage_model') is not None:
        mt = po['stochastic_power_outage_model']['type']
        if mt not in _OUTAGE_MODELS:
            raise UnsupportedSchemaError(f"power outage model '{mt}' is not built in")
        model = _OUTAGE_MODELS[mt](**(po['stochastic_power_outage_model'].get('attributes', {}) or {}))
    else:
        model = PowerOutage()

    # ---- devices (citylearn/citylearn.py:2326-2404) ----
    solar_generation = kwargs.get('solar_generation')
    solar_generation = True if solar_generation is None else solar_generation
    solar_generation = solar_generation[index] if isinstance(solar_generation, list) else solar_generation
    devices: Dict[str, Dict[str, Any]] = {}
    for dn in ('cooling_device', 'heating_device', 'dhw_device', 'dhw_storage', 'cooling_storage', 'heating_storage',
               'electrical_storage', 'pv'):
        ds = bs.get(dn)
        if ds is None or (dn == 'pv' and not solar_generation):
            continue
        if ds.get('autosize') and dn in ('electrical_storage', 'pv'):
            # Battery / PV autosizing samples external sizing tables and runs SAM's PVWatts (citylearn/energy_model.py:490-600,
            # 1143-1260): a dataset-construction step, not part of the stepped path.
            raise UnsupportedSchemaError(f"building '{name}': autosize of '{dn}' is outside the accelerated hot path (SURVEY.md §2 row 2e)")
        dt = ds['type']
        attrs = dict(ds.get('attributes', {}) or {})
        seed = attrs.pop('random_seed', None)
        seed = device_random_seed(name, building_type, dn, dt, schema_seed) if seed is None else seed
        if dn == 'pv':
            devices[dn] = {'type': 'PV', 'nominal_power': 0.0 if attrs.get('nominal_power') is None else float(attrs['nominal_power'])}
            continue
        if dt not in _DEVICE_RESOLVERS:
            raise UnsupportedSchemaError(f"device type '{dt}' is not supported")
        attrs.pop('seconds_per_time_step', None)
        devices[dn] = _DEVICE_RESOLVERS[dt](attrs, seed)
        devices[dn]['class'] = dt
        if ds.get('autosize'):
            _autosize(dn, devices[dn], series, spec, seed, dict(ds.get('autosize_attributes') or {}))
    for dn, resolver in _ABSENT_DEFAULT.items():
        if dn not in devices:
            if dn == 'electrical_storage':
                devices[dn] = resolve_battery({'capacity': 0.0, 'nominal_power': 0.0}, _absent_seed(name, dn))
            elif resolver is resolve_storage_tank:
                devices[dn] = resolver({'capacity': 0.0}, _absent_seed(name, dn))
            else:
                devices[dn] = resolver({'nominal_power': 0.0}, _absent_seed(name, dn))
            devices[dn]['absent'] = True
    devices.setdefault('pv', {'type': 'PV', 'nominal_power': 0.0, 'absent': True})
    if devices['cooling_device']['type'] != 'HeatPump':
        raise UnsupportedSchemaError('cooling_device must be a HeatPump')

    # ---- dynamics ----
    dyn_attrs, dyn_weights = None, None
    is_dyn = SUPPORTED_BUILDING_TYPES[building_type]
    if is_dyn:
        d = bs.get('dynamics')
        if d is None or d['type'] != 'citylearn.dynamics.LSTMDynamics':
            raise UnsupportedSchemaError('DynamicsBuilding needs citylearn.dynamics.LSTMDynamics')
        dyn_attrs = dict(d.get('attributes', {}))
        dyn_weights = source.state_dict(dyn_attrs['filename'])

    b = BuildingSpec(
        name=name, index=index, building_type=building_type, dynamics=is_dyn, observation_metadata=om, action_metadata=am,
        series=series, devices=devices, time_step_ratio=1.0 if time_step_ratio is None else float(time_step_ratio),
        seconds_per_time_step=seconds_per_time_step, simulate_power_outage=bool(simulate), stochastic_power_outage=bool(stochastic),
        outage_model=model, dynamics_attrs=dyn_attrs, dynamics_weights=dyn_weights, chargers=chargers, washing_machines=wms,
        charging_constraints=cc, observation_value_order=value_order)
    if cc is not None:
        for k, v in cc.one_hot.items():
            series[k] = np.full(n, v, dtype='float32')
    for w in wms:      # observation columns (citylearn/building.py:1298-1335): the machine's schedule at the observed time step
        series[f'{w.name}_start_time_step'] = np.asarray(w.start, dtype='float32')
        series[f'{w.name}_end_time_step'] = np.asarray(w.end, dtype='float32')
    if 'cooling_or_heating_device' in b.active_actions:
        assert 'cooling_device' not in b.active_actions and 'heating_device' not in b.active_actions, \
            'cooling_device and heating_device actions must be set to False when cooling_or_heating_device is True.'
    else:
        assert not ('cooling_device' in b.active_actions and 'heating_device' in b.active_actions), \
            'cooling_device and heating_device actions cannot both be set to True.'
    lo, hi = estimate_observation_space_limits(b, spec, include_all=False, periodic_normalization=False)
    b.observation_low = np.array(list(lo.values()), dtype='float32')
    b.observation_high = np.array(list(hi.values()), dtype='float32')
    b.action_low, b.action_high = estimate_action_space(b, spec)
    return b


# ------------------------------------------------------------------------------------------------
# spaces (citylearn/building.py:1836-2106, 2161-2282)
# ------------------------------------------------------------------------------------------------
def _window(a: np.ndarray, spec: DistrictSpec) -> np.ndarray:
    return a[spec.simulation_start_time_step:spec.simulation_end_time_step + 1]


def pv_generation(b: BuildingSpec, series: np.ndarray) -> np.ndarray:
    """`PV.get_generation` (citylearn/energy_model.py:469-488): python float * float64(series) / 1000."""
    return b.devices['pv']['nominal_power'] * np.array(series, dtype='float64') / 1000.0


def _bmin(a):
    """Python's `min(array)` as the reference uses it (citylearn/building.py:1836-2106), vectorised when no NaN can change the answer."""
    a = np.asarray(a)
    return min(a) if a.ndim != 1 or a.size == 0 or np.isnan(a).any() else a[np.argmin(a)]


def _bmax(a):
    a = np.asarray(a)
    return max(a) if a.ndim != 1 or a.size == 0 or np.isnan(a).any() else a[np.argmax(a)]


def estimate_observation_space_limits(b: BuildingSpec, spec: DistrictSpec, include_all=False, periodic_normalization=False):
    internal = ['net_electricity_consumption_without_storage', 'net_electricity_consumption_without_storage_and_partial_load',
                'net_electricity_consumption_without_storage_and_partial_load_and_pv']
    names = list(b.observation_metadata.keys()) + internal if include_all else b.active_observations
    dv = b.devices
    data = {k: _window(v, spec) for k, v in b.series.items() if v is not None}
    data['solar_generation'] = np.array(pv_generation(b, _window(b.series['solar_generation'], spec)))
    t_out = data['outdoor_dry_bulb_temperature']
    low: Dict[str, float] = {}
    high: Dict[str, float] = {}

    def dev_eff(d, heating):
        if d['type'] == 'HeatPump':
            cop = cop32(d, t_out, heating)
            return _bmin(cop), _bmax(cop)
        return d['efficiency'], d['efficiency']

    def input_power(d, demand, heating):
        if d['type'] == 'HeatPump':
            return demand / cop32(d, t_out, heating)
        return np.array(demand) / d['efficiency']

    cc = b.charging_constraints
    if cc is not None:       # building.py:2138-2157
        for k in cc.headroom_names:
            lim = cc.building_limit_kw if k == 'charging_building_headroom_kw' else next(
                ph['limit_kw'] for ph in cc.phases if k == f"charging_phase_{ph['name']}_headroom_kw")
            data[k] = np.full(2, float(lim), dtype='float32')
    for key in names:
        if key.startswith('charging_phase_one_hot_'):
            low[key], high[key] = 0.0, 1.0
        elif key == 'charging_constraint_violation_kwh':
            low[key] = 0.0
            high[key] = sum((c.max_charging_power or 0.0) for c in b.chargers) * (b.seconds_per_time_step / 3600)
        elif key == 'net_electricity_consumption':
            lows = data['non_shiftable_load'] - (+dv['electrical_storage']['nominal_power'] + data['solar_generation'])
            highs = (data['non_shiftable_load'] + dv['cooling_device']['nominal_power'] + dv['heating_device']['nominal_power']
                     + dv['dhw_device']['nominal_power'] + dv['electrical_storage']['nominal_power'] - data['solar_generation'])
            low[key] = min(lows.min(), 0.0)
            high[key] = highs.max()
        elif key == 'net_electricity_consumption_without_storage':
            low[key] = min(low['net_electricity_consumption'] + dv['electrical_storage']['nominal_power'], 0.0)
            high[key] = high['net_electricity_consumption'] - dv['electrical_storage']['nominal_power']
        elif key == 'net_electricity_consumption_without_storage_and_partial_load':
            low[key] = low['net_electricity_consumption_without_storage']
            high[key] = high['net_electricity_consumption_without_storage']
        elif key == 'net_electricity_consumption_without_storage_and_partial_load_and_pv':
            low[key] = 0.0
            high[key] = (data['non_shiftable_load'] + dv['cooling_device']['nominal_power'] + dv['heating_device']['nominal_power']
                         + dv['dhw_device']['nominal_power']).max()
        elif key in ('cooling_storage_soc', 'heating_storage_soc', 'dhw_storage_soc', 'electrical_storage_soc'):
            low[key], high[key] = 0.0, 1.0
        elif key == 'cooling_device_efficiency':
            low[key], high[key] = dev_eff(dv['cooling_device'], False)
        elif key == 'heating_device_efficiency':
            low[key], high[key] = dev_eff(dv['heating_device'], True)
        elif key == 'dhw_device_efficiency':
            low[key], high[key] = dev_eff(dv['dhw_device'], True)
        elif key == 'indoor_dry_bulb_temperature':
            low[key] = data[key].min() - b.maximum_temperature_delta
            high[key] = data[key].max() + b.maximum_temperature_delta
        elif key in ('indoor_dry_bulb_temperature_cooling_delta', 'indoor_dry_bulb_temperature_heating_delta'):
            low[key], high[key] = -b.maximum_temperature_delta, b.maximum_temperature_delta
        elif key == 'comfort_band':
            low[key], high[key] = 0, _bmax(data[key])
        elif key in ('cooling_demand', 'heating_demand', 'dhw_demand'):
            low[key] = 0.0
            high[key] = data[key].max() * b.demand_observation_limit_factor
        elif key == 'cooling_electricity_consumption':
            low[key], high[key] = 0.0, dv['cooling_device']['nominal_power']
        elif key == 'heating_electricity_consumption':
            low[key], high[key] = 0.0, dv['heating_device']['nominal_power']
        elif key == 'dhw_electricity_consumption':
            low[key], high[key] = 0.0, dv['dhw_device']['nominal_power']
        elif key == 'cooling_storage_electricity_consumption':
            low[key] = -_bmax(input_power(dv['cooling_device'], data['cooling_demand'], False))
            high[key] = dv['cooling_device']['nominal_power']
        elif key == 'heating_storage_electricity_consumption':
            low[key] = -_bmax(input_power(dv['heating_device'], data['heating_demand'], True))
            high[key] = dv['heating_device']['nominal_power']
        elif key == 'dhw_storage_electricity_consumption':
            low[key] = -_bmax(input_power(dv['dhw_device'], data['dhw_demand'], True))
            high[key] = dv['dhw_device']['nominal_power']
        elif key == 'electrical_storage_electricity_consumption':
            low[key], high[key] = -dv['electrical_storage']['nominal_power'], dv['electrical_storage']['nominal_power']
        elif key == 'power_outage':
            low[key], high[key] = 0.0, 1.0
        elif periodic_normalization and key in PERIODIC:
            x = 2 * np.pi * np.array(list(range(1, PERIODIC[key] + 1))) / PERIODIC[key]
            x_sin, x_cos = np.sin(x), np.cos(x)
            low[f'{key}_cos'], high[f'{key}_cos'] = _bmin(x_cos), _bmax(x_cos)
            low[f'{key}_sin'], high[f'{key}_sin'] = _bmin(x_sin), _bmax(x_sin)
        elif 'connected_state' in key or '_incoming_state' in key:
            low[key], high[key] = 0, 1
        elif '_departure_time' in key or '_estimated_arrival_time' in key:
            low[key], high[key] = -1, 24
        elif '_soc' in key and '_electric_vehicle' in key:
            low[key], high[key] = -0.1, 1.0
        elif 'charger' in key and key.endswith('_battery_capacity'):
            low[key], high[key] = -1, 100
        elif any(key in (f'{w.name}_start_time_step', f'{w.name}_end_time_step') for w in b.washing_machines):
            low[key], high[key] = -1, 24
        elif key == 'occupant_interaction_indoor_dry_bulb_temperature_set_point_delta':
            pass
        else:
            if key not in data:
                raise UnsupportedSchemaError(f"observation '{key}' is outside the accelerated hot path")
            low[key], high[key] = _bmin(data[key]), _bmax(data[key])
    d = b.observation_space_limit_delta
    return {k: v - d for k, v in low.items()}, {k: v + d for k, v in high.items()}


def estimate_action_space(b: BuildingSpec, spec: DistrictSpec) -> Tuple[np.ndarray, np.ndarray]:
    lo, hi = [], []
    dv = b.devices
    for key in b.active_actions:
        if key == 'cooling_or_heating_device':
            lo.append(-1.0 if dv['cooling_device']['nominal_power'] > EPS else 0.0)
            hi.append(1.0 if dv['heating_device']['nominal_power'] > EPS else 0.0)
        elif key in ('cooling_device', 'heating_device'):
            lo.append(0.0)
            hi.append(1.0)
        elif key == 'electrical_storage':
            lo.append(-1.0)
            hi.append(1.0)
        elif key in ('cooling_storage', 'heating_storage', 'dhw_storage'):
            end_use = key.split('_')[0]
            limit = dv[f'{end_use}_device']['nominal_power'] / max(dv[key]['capacity'], EPS)
            limit = min(limit, 1.0)
            lo.append(-limit)
            hi.append(limit)
        elif key.startswith('electric_vehicle_storage_'):       # citylearn/building.py:2199-2205
            c = next(c for c in b.chargers if key == f'electric_vehicle_storage_{c.charger_id}')
            lo.append(0.0 if c.max_discharging_power == 0 else -1.0)
            hi.append(1.0)
        elif any(key == w.name for w in b.washing_machines):   # :2207-2212
            lo.append(0.0)
            hi.append(1.0)
        else:
            raise UnsupportedSchemaError(f"action '{key}' is outside the accelerated hot path")
    return np.array(lo, dtype='float32'), np.array(hi, dtype='float32')


# ------------------------------------------------------------------------------------------------
# finalize: flat arrays for the device
# ------------------------------------------------------------------------------------------------
def _observation_source(b: BuildingSpec, name: str):
    """(kind, payload) of one observation name for building b; series names resolve to table columns."""
    if name in DYN and name not in ('cooling_demand', 'heating_demand', 'dhw_demand', 'indoor_dry_bulb_temperature'):
        return ('dyn', name)
    if name == 'power_outage':
        return ('outage', None)
    if b.charging_constraints is not None and (name in b.charging_constraints.headroom_names or name == 'charging_constraint_violation_kwh'):
        return ('state', name)
    if name == 'solar_generation':
        return ('derived', 'solar_generation_obs')
    if name in ('cooling_device_efficiency', 'heating_device_efficiency', 'dhw_device_efficiency',
                'indoor_dry_bulb_temperature_cooling_delta', 'indoor_dry_bulb_temperature_heating_delta'):
        return ('derived', name)
    if name in b.series and b.series[name] is not None:
        return ('series', name)
    raise UnsupportedSchemaError(f"observation '{name}' is outside the accelerated hot path")


def derived_series(b: BuildingSpec, name: str) -> np.ndarray:
    s, dv = b.series, b.devices
    if name == 'solar_neg':      # building.py:2554  (float64, stored float32)
        return (pv_generation(b, s['solar_generation']) * -1).astype('float32')
    if name == 'solar_generation_obs':   # building.py:1428
        return np.abs(pv_generation(b, s['solar_generation']) * -1).astype('float32')
    if name == 'cooling_device_efficiency':
        return cop32(dv['cooling_device'], s['outdoor_dry_bulb_temperature'], False).astype('float32')
    if name in ('heating_device_efficiency', 'dhw_device_efficiency'):
        d = dv['heating_device' if name.startswith('heating') else 'dhw_device']
        if d['type'] == 'HeatPump':
            return cop32(d, s['outdoor_dry_bulb_temperature'], True).astype('float32')
        return np.full(len(s['hour']), d['efficiency'], dtype='float32')
    if name == 'indoor_dry_bulb_temperature_cooling_delta':
        return (s['indoor_dry_bulb_temperature'] - s['indoor_dry_bulb_temperature_cooling_set_point']).astype('float32')
    if name == 'indoor_dry_bulb_temperature_heating_delta':
        return (s['indoor_dry_bulb_temperature'] - s['indoor_dry_bulb_temperature_heating_set_point']).astype('float32')
    raise KeyError(name)


def lstm_exogenous_inputs(b: BuildingSpec) -> Tuple[List[np.ndarray], int, int]:
    """Pre-normalised LSTM inputs that do not depend on actions (citylearn/building.py:3057-3078, 1191-1204).

    Returns (columns in input order with None at the two fed-back slots, slot of indoor temperature, slot of cooling demand).
    """
    a = b.dynamics_attrs
    cols: List[Optional[np.ndarray]] = []
    slot_tin = slot_cdem = -1
    for i, (k, mn, mx) in enumerate(zip(a['input_observation_names'], a['input_normalization_minimum'], a['input_normalization_maximum'])):
        if k == 'indoor_dry_bulb_temperature':
            slot_tin = i
            cols.append(None)
            continue
        if k == 'cooling_demand':
            slot_cdem = i
            cols.append(None)
            continue
        if k.endswith('_sin') or k.endswith('_cos'):
            base = k[:-4]
            x = 2 * np.pi * b.series[base] / PERIODIC[base]          # int32 -> float64 (preprocessing.py:68-72)
            v = np.sin(x) if k.endswith('_sin') else np.cos(x)
            cols.append(((v - mn) / (mx - mn)).astype('float32'))
        elif k in b.series and b.series[k] is not None:
            v = b.series[k]
            cols.append(((v - mn) / (mx - mn)).astype('float32'))   # float32 series: weak-scalar float32 arithmetic
        else:
            raise UnsupportedSchemaError(f"LSTM input '{k}' is not supported")
    if slot_tin < 0:
        raise UnsupportedSchemaError('LSTM dynamics without indoor_dry_bulb_temperature input')
    return cols, slot_tin, slot_cdem


def finalize(spec: DistrictSpec) -> None:
    """Build the time-major table, the per-building parameter blocks and the action layout."""
    B = spec.n_buildings
    n = len(spec.buildings[0].series['hour'])
    for b in spec.buildings:
        assert len(b.series['hour']) == n, 'all buildings must share the dataset length'
    cols: List[np.ndarray] = []
    index: Dict[Any, int] = {}
    dedup: Dict[bytes, int] = {}

    def add(key, arr: np.ndarray) -> int:
        a32 = np.ascontiguousarray(arr, dtype='float32')
        h = hashlib.sha1(a32.tobytes()).digest()
        if h in dedup:                       # identical series (shared weather/pricing files) stored once
            index[key] = dedup[h]
        else:
            dedup[h] = len(cols)
            index[key] = len(cols)
            cols.append(a32)
        return index[key]

    params = np.zeros((B, NPARAM), dtype='float64')
    iparams = np.full((B, NIPARAM), -1, dtype='int32')
    weights: List[np.ndarray] = []
    w_off = 0
    a_off = 0
    ev_action_slot: Dict[Any, int] = {}
    wm_action_slot: Dict[Any, int] = {}
    for bi, b in enumerate(spec.buildings):
        s, dv = b.series, b.devices
        ip = iparams[bi]
        p = params[bi]
        physics = {
            'C_NSL': s['non_shiftable_load'], 'C_DHW_DEMAND': s['dhw_demand'], 'C_COOLING_DEMAND': s['cooling_demand'],
            'C_HEATING_DEMAND': s['heating_demand'], 'C_SOLAR': s['solar_generation'],
            'C_T_OUT': s['outdoor_dry_bulb_temperature'], 'C_PRICE': s['electricity_pricing'], 'C_CARBON': s['carbon_intensity'],
            'C_HVAC_MODE': s['hvac_mode'], 'C_T_IN': s['indoor_dry_bulb_temperature'],
            'C_COOL_SP': s['indoor_dry_bulb_temperature_cooling_set_point'],
            'C_HEAT_SP': s['indoor_dry_bulb_temperature_heating_set_point'], 'C_COMFORT_BAND': s['comfort_band'],
            'C_OCCUPANT': s['occupant_count'],
        }
        for k, v in physics.items():
            ip[IP[k]] = add((bi, k), v)
        for name in b.observation_metadata:
            try:
                kind, payload = _observation_source(b, name)
            except UnsupportedSchemaError:
                if b.observation_metadata[name]:
                    raise
                continue
            if kind == 'series':
                add((bi, name), s[name])
            elif kind == 'derived':
                add((bi, name), derived_series(b, payload))
        for name in [k for k in s if k.endswith('__t0')]:          # episode-start variants of observation columns (chargers)
            add((bi, name), s[name])
        # dyn-mapped observation names that are table-backed in reference-parity ("stale") mode
        for name in ('cooling_demand', 'heating_demand', 'dhw_demand', 'indoor_dry_bulb_temperature'):
            add((bi, name), s[name])

        bat = dv['electrical_storage']
        p[P['BAT_CAPACITY']] = bat['capacity']
        p[P['BAT_NOMINAL_POWER']] = bat['nominal_power']
        p[P['BAT_EFFICIENCY0']] = bat['efficiency']
        p[P['BAT_LOSS']] = bat['loss_coefficient']
        p[P['BAT_CLC']] = bat['capacity_loss_coefficient']
        p[P['BAT_DOD']] = bat['depth_of_discharge']
        p[P['BAT_INITIAL_SOC']] = bat['initial_soc']
        pe, cp = bat['power_efficiency_curve'], bat['capacity_power_curve']
        ip[IP['PE_N']], ip[IP['CP_N']] = pe.shape[1], cp.shape[1]
        for j in range(MAX_CURVE):
            p[P['PE_X0'] + j] = pe[0][min(j, pe.shape[1] - 1)]
            p[P['PE_Y0'] + j] = pe[1][min(j, pe.shape[1] - 1)]
            p[P['CP_X0'] + j] = cp[0][min(j, cp.shape[1] - 1)]
            p[P['CP_Y0'] + j] = cp[1][min(j, cp.shape[1] - 1)]
        flags = 0
        for pre, dn in (('CS', 'cooling_storage'), ('HS', 'heating_storage'), ('DS', 'dhw_storage')):
            t = dv[dn]
            p[P[f'{pre}_CAPACITY']] = t['capacity']
            if t.get('autosized') and pre != 'DS':
                flags |= F_CS_CAPACITY_F32 if pre == 'CS' else F_HS_CAPACITY_F32
            p[P[f'{pre}_EFFICIENCY']] = t['efficiency']
            p[P[f'{pre}_LOSS']] = t['loss_coefficient']
            p[P[f'{pre}_INITIAL_SOC']] = t['initial_soc']
            if t.get('max_input_power') is not None:
                p[P[f'{pre}_MAX_IN']] = t['max_input_power']
                flags |= {'CS': F_CS_HAS_MAX_IN, 'HS': F_HS_HAS_MAX_IN, 'DS': F_DS_HAS_MAX_IN}[pre]
            if t.get('max_output_power') is not None:
                p[P[f'{pre}_MAX_OUT']] = t['max_output_power']
                flags |= {'CS': F_CS_HAS_MAX_OUT, 'HS': F_HS_HAS_MAX_OUT, 'DS': F_DS_HAS_MAX_OUT}[pre]
        cd = dv['cooling_device']
        p[P['CD_NOMINAL_POWER']] = cd['nominal_power']
        if cd.get('autosized'):
            flags |= F_CD_NOMINAL_F32
        if dv['heating_device'].get('autosized'):
            flags |= F_HD_NOMINAL_F32
        p[P['CD_COP_NUM']] = cd['efficiency'] * (cd['target_cooling_temperature'] + 273.15)
        p[P['CD_TARGET']] = cd['target_cooling_temperature']
        for pre, dn, fl in (('HD', 'heating_device', F_HEATING_IS_HEAT_PUMP), ('DD', 'dhw_device', F_DHW_IS_HEAT_PUMP)):
            d = dv[dn]
            p[P[f'{pre}_NOMINAL_POWER']] = d['nominal_power']
            if d['type'] == 'HeatPump':
                flags |= fl
                p[P[f'{pre}_COP_NUM']] = d['efficiency'] * (d['target_heating_temperature'] + 273.15)
                p[P[f'{pre}_TARGET']] = d['target_heating_temperature']
                p[P[f'{pre}_EFFICIENCY']] = 1.0
            else:
                p[P[f'{pre}_EFFICIENCY']] = d['efficiency']
        # quirk kept: at t == 0 a heater-type heating device is billed with the DHW heater's efficiency (building.py:2632)
        p[P['TIME_STEP_RATIO']] = b.time_step_ratio
        p[P['PV_NOMINAL_POWER']] = dv['pv']['nominal_power']
        p[P['HOURS_PER_STEP']] = spec.seconds_per_time_step / 3600
        thermal = (s['cooling_demand'].any() or s['heating_demand'].any() or s['dhw_demand'].any()
                   or any(dv[k]['capacity'] > 0 for k in ('cooling_storage', 'heating_storage', 'dhw_storage'))
                   or any(dv[k]['nominal_power'] > 0 for k in ('cooling_device', 'heating_device', 'dhw_device'))
                   or b.dynamics)
        if thermal:
            flags |= F_HAS_THERMAL
        if b.simulate_power_outage:
            flags |= F_SIMULATE_OUTAGE
        if b.dynamics:
            flags |= F_DYNAMICS
            a = b.dynamics_attrs
            ex, slot_tin, slot_cdem = lstm_exogenous_inputs(b)
            first = None
            for i, c in enumerate(ex):
                # fed-back slots get a placeholder column so that the input columns stay contiguous
                col = np.zeros(n, dtype='float32') if c is None else c
                cols.append(np.ascontiguousarray(col, dtype='float32'))   # no dedup: contiguity matters
                index[(bi, 'lstm_in', i)] = len(cols) - 1
                first = len(cols) - 1 if first is None else first
            ip[IP['DYN_C_INPUTS']] = first
            ip[IP['DYN_N_INPUTS']] = len(ex)
            ip[IP['DYN_SLOT_TIN']] = slot_tin
            ip[IP['DYN_SLOT_CDEM']] = slot_cdem
            ip[IP['DYN_LOOKBACK']] = int(a['lookback'])
            ip[IP['DYN_HIDDEN']] = int(a['hidden_size'])
            if int(a['num_layers']) != 2:
                raise UnsupportedSchemaError('LSTM dynamics: only num_layers == 2 is supported')
            p[P['DYN_TIN_MIN']] = a['input_normalization_minimum'][slot_tin]
            p[P['DYN_TIN_MAX']] = a['input_normalization_maximum'][slot_tin]
            if slot_cdem >= 0:
                p[P['DYN_CDEM_MIN']] = a['input_normalization_minimum'][slot_cdem]
                p[P['DYN_CDEM_MAX']] = a['input_normalization_maximum'][slot_cdem]
            w = b.dynamics_weights
            blk = np.concatenate([
                w['l_lstm.weight_ih_l0'].ravel(), w['l_lstm.weight_hh_l0'].ravel(),
                w['l_lstm.bias_ih_l0'].ravel(), w['l_lstm.bias_hh_l0'].ravel(),
                w['l_lstm.weight_ih_l1'].ravel(), w['l_lstm.weight_hh_l1'].ravel(),
                w['l_lstm.bias_ih_l1'].ravel(), w['l_lstm.bias_hh_l1'].ravel(),
                w['l_linear.weight'].ravel(), w['l_linear.bias'].ravel(),
            ]).astype('float32')
            ip[IP['DYN_W_OFFSET']] = w_off
            weights.append(blk)
            w_off += blk.size
        ip[IP['FLAGS']] = flags
        # action slots inside the district action vector (building order, active actions in schema order)
        for an in b.active_actions:
            if an.startswith('electric_vehicle_storage_'):
                ev_action_slot[(bi, an[len('electric_vehicle_storage_'):])] = a_off
            elif any(an == w.name for w in b.washing_machines):
                wm_action_slot[(bi, an)] = a_off
            else:
                ip[IP['A_' + an.upper()]] = a_off
            a_off += 1
    spec.action_dim = a_off
    # ---- electric vehicles / chargers / washing machines: flat arrays for the device and the oracle ----
    chargers = [c for b in spec.buildings for c in b.chargers]
    wms = [w for b in spec.buildings for w in b.washing_machines]
    if chargers or wms or spec.evs:
        from . import ev as EV
        sched = (spec.ev or {}).get('schedule')
        if sched is None:
            sched = EV.compile_schedule(chargers, len(spec.evs), n, spec.simulation_start_time_step, spec.simulation_end_time_step, spec.ev_random_seed)
        n_ev = len(spec.evs)
        ev_params = np.zeros((n_ev, NPARAM), dtype='float64')
        ev_ip = np.zeros((n_ev, 2), dtype='int32')
        ev_cols = np.zeros((n_ev, 4), dtype='int32')
        for v, e in enumerate(spec.evs):
            bat, q = e.battery, ev_params[v]
            q[P['BAT_CAPACITY']], q[P['BAT_NOMINAL_POWER']], q[P['BAT_EFFICIENCY0']] = bat['capacity'], bat['nominal_power'], bat['efficiency']
            q[P['BAT_LOSS']], q[P['BAT_CLC']], q[P['BAT_DOD']], q[P['BAT_INITIAL_SOC']] = bat['loss_coefficient'], bat['capacity_loss_coefficient'], bat['depth_of_discharge'], bat['initial_soc']
            pe, cp = bat['power_efficiency_curve'], bat['capacity_power_curve']
            ev_ip[v] = (pe.shape[1], cp.shape[1])
            for j in range(MAX_CURVE):
                q[P['PE_X0'] + j] = pe[0][min(j, pe.shape[1] - 1)]; q[P['PE_Y0'] + j] = pe[1][min(j, pe.shape[1] - 1)]
                q[P['CP_X0'] + j] = cp[0][min(j, cp.shape[1] - 1)]; q[P['CP_Y0'] + j] = cp[1][min(j, cp.shape[1] - 1)]
            q[P['TIME_STEP_RATIO']] = spec.buildings[0].time_step_ratio
            q[P['HOURS_PER_STEP']] = spec.seconds_per_time_step / 3600
            for j, key in enumerate(('assoc', 'sim', 'drift', 't0')):     # NaN = not applicable on that row
                cols.append(np.ascontiguousarray(sched[key][:, v], dtype='float32'))
                index[('ev', v, key)] = len(cols) - 1
                ev_cols[v, j] = len(cols) - 1
        ch_building = np.array([c.building for c in chargers], dtype='int32')
        ch_action = np.array([ev_action_slot.get((c.building, c.charger_id), -1) for c in chargers], dtype='int32')
        ch_cols = np.zeros((len(chargers), 4), dtype='int32')
        CHP = EV.CHARGER_PARAMS
        ch_params = np.zeros((len(chargers), len(CHP)), dtype='float64')
        for k, c in enumerate(chargers):
            con = EV.connected_mask(c)
            series = {'conn': con.astype('float32'), 'ev': np.where(con, c.ev, -1).astype('float32'),
                      'req': np.asarray(c.required_soc, dtype='float32'), 'dep': np.asarray(c.departure_time, dtype='float32')}
            for j, key in enumerate(('conn', 'ev', 'req', 'dep')):
                cols.append(np.ascontiguousarray(series[key]))
                index[('charger', c.building, c.charger_id, key)] = len(cols) - 1
                ch_cols[k, j] = len(cols) - 1
            q = ch_params[k]
            q[CHP['MAX_C']], q[CHP['MIN_C']], q[CHP['MAX_D']], q[CHP['MIN_D']], q[CHP['EFF']] = (c.max_charging_power, c.min_charging_power,
                                                                                                 c.max_discharging_power, c.min_discharging_power, c.efficiency)
            for pre, cv in (('C', c.charge_curve), ('D', c.discharge_curve)):
                if cv is None:
                    continue
                if cv.shape[1] > MAX_CURVE:
                    raise UnsupportedSchemaError('charger efficiency curves with more than 8 points are not supported')
                q[CHP[f'{pre}_N']] = cv.shape[1]
                for j in range(cv.shape[1]):
                    q[CHP[f'{pre}_X0'] + j], q[CHP[f'{pre}_Y0'] + j] = cv[0][j], cv[1][j]
        wm_building = np.array([w.building for w in wms], dtype='int32')
        wm_action = np.array([wm_action_slot.get((w.building, w.name), -1) for w in wms], dtype='int32')
        wm_cols = np.zeros((len(wms), 4), dtype='int32')
        for k, w in enumerate(wms):
            # the load column is FOLLOWED by the partial-sum columns (first 1, 2, ... entries of the profile): the device reads
            # `load + (T - t)` for a cycle whose profile would run past the episode end (include/citylearn_b200.h, cl_ev_desc.wm_cols)
            for j, (key, arr) in ((0, ('start', w.start)), (1, ('end', w.end)), (3, ('len', w.profile_len)), (2, ('load', w.profile_sum))):
                cols.append(np.ascontiguousarray(arr, dtype='float32'))
                index[('wm', w.building, w.name, key)] = len(cols) - 1
                wm_cols[k, j] = len(cols) - 1
            for j in range(w.profile_prefix.shape[1]):
                cols.append(np.ascontiguousarray(w.profile_prefix[:, j], dtype='float32'))
                index[('wm', w.building, w.name, f'load_first_{j + 1}')] = len(cols) - 1
        # charging constraints (building.py:764-989): limits (NaN: none), phase member lists as indices into the building's chargers
        cc_buildings = [bi for bi, b in enumerate(spec.buildings) if b.charging_constraints is not None]
        cc_limits = np.full((len(cc_buildings), 1 + EV.MAX_PHASES), np.nan, dtype='float64')
        cc_members = np.full((len(cc_buildings), EV.MAX_PHASES, 4), -1, dtype='int32')
        for k, bi in enumerate(cc_buildings):
            b, cc = spec.buildings[bi], spec.buildings[bi].charging_constraints
            if not b.chargers:
                continue
            ids = [c.charger_id for c in b.chargers]
            if cc.building_limit_kw is not None:
                cc_limits[k, 0] = float(cc.building_limit_kw)
            for j, ph in enumerate(cc.phases):
                if ph['limit_kw'] is not None:
                    cc_limits[k, 1 + j] = float(ph['limit_kw'])
                mem = [ids.index(cid) for cid in ph['chargers'] if cid in ids]
                if len(mem) > 4:
                    raise UnsupportedSchemaError('a charging phase lists more than 4 charger entries')
                cc_members[k, j, :len(mem)] = mem
        spec.ev = {'schedule': sched, 'n_ev': n_ev, 'ev_params': ev_params, 'ev_ip': ev_ip, 'ev_cols': ev_cols, 'chargers': chargers,
                   'ch_building': ch_building, 'ch_action': ch_action, 'ch_cols': ch_cols, 'ch_params': ch_params, 'wms': wms,
                   'wm_building': wm_building, 'wm_action': wm_action, 'wm_cols': wm_cols,
                   'cc_building': np.array(cc_buildings, dtype='int32'), 'cc_limits': cc_limits, 'cc_members': cc_members}
    spec.table = np.ascontiguousarray(np.stack(cols, axis=1), dtype='float32')
    spec.columns = index
    spec.params = params
    spec.iparams = iparams
    spec.lstm_weights = np.concatenate(weights) if weights else np.zeros(0, dtype='float32')


# ------------------------------------------------------------------------------------------------
# observation layout
# ------------------------------------------------------------------------------------------------
def observation_layout(spec: DistrictSpec, central_agent: Optional[bool] = None, stale: bool = True, names_order: bool = False):
    """Flat observation row of one env: list of (building index, name) and the matching device descriptors.

    Decentralised: concatenation of every building's active observations (citylearn/citylearn.py:482-483).
    Central agent: building 0 complete, later buildings drop shared names already seen (citylearn/citylearn.py:462-480).

    Descriptor rows are int32 [L, 4] = (kind, a, b, building).  With `stale=True` (reference parity)
    demand / indoor-temperature observations read the dataset column; action-dependent ones are DYN slots which the
    step kernel zeroes after a step and fills at reset.

    The row follows the order of the VALUES `Building.observations()` returns; `names_order=True` gives the entries in the order of the
    NAMES (`active_observations`, the order of the observation space) instead - they differ for buildings with charging constraints,
    where the reference itself is inconsistent (building.py:1146-1154 vs :811-833).
    """
    central = spec.central_agent if central_agent is None else central_agent
    entries: List[Tuple[int, str]] = []
    seen: List[str] = []
    for bi, b in enumerate(spec.buildings):
        for name in (names_order and b.active_observations) or b.observation_value_order or b.active_observations:
            if (not central) or bi == 0 or name not in spec.shared_observations or name not in seen:
                entries.append((bi, name))
            if central and name in spec.shared_observations and name not in seen:
                seen.append(name)
    if names_order:
        return entries, None
    desc = np.zeros((len(entries), 4), dtype='int32')
    cc_index = {int(bi): k for k, bi in enumerate((spec.ev or {}).get('cc_building', []))}
    for j, (bi, name) in enumerate(entries):
        b = spec.buildings[bi]
        stale_ts = name in ('cooling_demand', 'heating_demand', 'dhw_demand', 'indoor_dry_bulb_temperature')
        cc = b.charging_constraints
        if cc is not None and name in cc.state_names:
            if name == 'charging_constraint_violation_kwh':
                which = CC_SLOTS - 1
            elif name == 'charging_building_headroom_kw':
                which = 0
            else:
                which = 1 + next(i for i, ph in enumerate(cc.phases) if name == f"charging_phase_{ph['name']}_headroom_kw")
            desc[j] = (OBS_STATE, cc_index[bi] * CC_SLOTS + which, 0, bi)
        elif name in DYN and not (stale and stale_ts):
            desc[j] = (OBS_DYN, DYN[name], 0, bi)
        elif name == 'power_outage':
            desc[j] = (OBS_OUTAGE, 0, 0, bi)
        else:
            alt = spec.columns.get((bi, name + '__t0'))       # value on the first row of an episode, when it differs (b = column + 1)
            desc[j] = (OBS_TS, spec.columns[(bi, name)], 0 if alt is None else alt + 1, bi)
    return entries, desc


PERIODIC_OBSERVATIONS = {'hour': 24, 'day_type': 7, 'month': 12, 'minutes': 60}   # max of the ranges in building.py:1484-1498
OBS_FN_IDENTITY, OBS_FN_SIN, OBS_FN_COS = 0, 1, 2
OBS_TRANSFORM_DTYPE = np.dtype([('fn', '<i4'), ('w', '<f4'), ('scale', '<f4'), ('offset', '<f4'), ('lo', '<f4'), ('hi', '<f4')])   # cl_obs_transform


def transformed_observation_layout(spec: DistrictSpec, entries, desc: np.ndarray, mode: str):
    """Observation row under a wrapper (`citylearn/wrappers.py:15-167`): (entries, descriptors, cl_obs_transform records).

    'normalized' (NormalizedObservationWrapper): every periodic observation (hour, day_type, month) becomes a `<name>_cos`,
    `<name>_sin` pair - in that order, `Building.observations(periodic_normalization=True)` building.py:1191-1204 - and every
    value is min-max scaled with the limits of `estimate_observation_space_limits(include_all=True, periodic_normalization=True)`
    (`x_min == x_max` gives 0, preprocessing.py:139-152).  'clipped' (ClippedObservationWrapper): values are clipped to the
    observation space.
    """
    out_entries, rows, recs = [], [], []
    limits = {}
    for j, (bi, name) in enumerate(entries):
        b = spec.buildings[bi]
        if mode == 'normalized':
            if bi not in limits:
                limits[bi] = estimate_observation_space_limits(b, spec, include_all=True, periodic_normalization=True)
            lo, hi = limits[bi]
            parts = ([(f'{name}_cos', OBS_FN_COS), (f'{name}_sin', OBS_FN_SIN)] if name in PERIODIC_OBSERVATIONS else [(name, OBS_FN_IDENTITY)])
            for n2, fn in parts:
                x_min, x_max = float(lo[n2]), float(hi[n2])
                scale, offset = (0.0, 0.0) if x_min == x_max else (1.0 / (x_max - x_min), -x_min / (x_max - x_min))
                w = 2.0 * math.pi / PERIODIC_OBSERVATIONS[name] if fn != OBS_FN_IDENTITY else 0.0
                out_entries.append((bi, n2))
                rows.append(desc[j])
                recs.append((fn, w, scale, offset, -np.inf, np.inf))
        elif mode == 'clipped':
            k = b.active_observations.index(name)
            out_entries.append((bi, name))
            rows.append(desc[j])
            recs.append((OBS_FN_IDENTITY, 0.0, 1.0, 0.0, float(b.observation_low[k]), float(b.observation_high[k])))
        else:
            raise ValueError(f"unknown observation transform '{mode}' (expected 'normalized' or 'clipped')")
    return out_entries, np.array(rows, dtype='int32').reshape(-1, 4), np.array(recs, dtype=OBS_TRANSFORM_DTYPE)


def outage_signals(spec: DistrictSpec, episode_time_steps: int, episode_start: int) -> np.ndarray:
    """[B, episode_time_steps] float32 0/1 (citylearn/building.py:2566-2594)."""
    out = np.zeros((spec.n_buildings, episode_time_steps), dtype='float32')
    for bi, b in enumerate(spec.buildings):
        if not b.simulate_power_outage:
            continue
        if b.stochastic_power_outage:
            sig = b.outage_model.get_signals(episode_time_steps, seconds_per_time_step=spec.seconds_per_time_step)
        else:
            sig = b.series['power_outage'][episode_start:episode_start + episode_time_steps]
        out[bi, :len(sig)] = np.asarray(sig, dtype='float32')[:episode_time_steps]
    return out
