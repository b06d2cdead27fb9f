"""Electric vehicles, chargers and washing machines: host-side ingestion and schedule pre-computation.

Reference: `citylearn/electric_vehicle_charger.py` (Charger), `citylearn/electric_vehicle.py` (ElectricVehicle),
`citylearn/energy_model.py:1244-1398` (WashingMachine), `citylearn/data.py:663-820` (ChargerSimulation / WashingMachineSimulation),
`citylearn/citylearn.py:1325-1475` (association of EVs to chargers, SOC of EVs that are away), `citylearn/building.py:1221-1335`
(charger / washing-machine observations), `:1536-1640` (actions), `:2615-2703` (consumption in `update_variables`).

What makes this path fit the step kernel: WHICH vehicle sits at WHICH charger at time step t, when it arrives, with which state of
charge, and what an away vehicle's battery does, all come from the dataset (plus one stream of random draws) - none of it depends on
the actions.  So the host compiles, per dataset row, (a) the charger's connection record, (b) every charger observation as a plain
table column (reference-parity observations read the zero-initialised / arrival value of `battery.soc[t]`) and (c) per
vehicle the operation that `next_time_step` applies to its `soc[t]` entry: nothing (0 stays), a constant (arrival SOC), or a factor on
`soc[t-1]` (the away-from-charger drift).  The device keeps per (vehicle, env): soc[t-1], soc[t], degraded capacity, efficiency and the
"has charged before" flag, and runs `Battery.charge` for connected vehicles inside the owning building's unit step.

Randomness.  The reference draws the drift factors from NumPy's GLOBAL generator (`np.random.normal(1.0, 0.2)`, citylearn.py:1473), one
draw per away vehicle per step in vehicle order, and a missing `initial_soc` from Python's global `random`.  Neither is reproducible
there unless the caller seeds the globals; here `ev_random_seed` (default: the schema's `random_seed`) seeds a private
`np.random.RandomState` whose legacy stream equals `np.random.seed(ev_random_seed)` - the fixtures are recorded that way.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Any, Dict, List, Mapping, Optional

import numpy as np

DEFAULT_TIME = -1          # citylearn/data.py:718-719
DEFAULT_SOC = -0.1

CHARGER_OBSERVATIONS = [   # per charger, in the order `process_metadata` appends them (citylearn/citylearn.py:2505-2540)
    ('electric_vehicle_charger_connected_state', 'electric_vehicle_charger_{id}_connected_state'),
    ('connected_electric_vehicle_at_charger_departure_time', 'connected_electric_vehicle_at_charger_{id}_departure_time'),
    ('connected_electric_vehicle_at_charger_required_soc_departure', 'connected_electric_vehicle_at_charger_{id}_required_soc_departure'),
    ('connected_electric_vehicle_at_charger_soc', 'connected_electric_vehicle_at_charger_{id}_soc'),
    ('connected_electric_vehicle_at_charger_battery_capacity', 'connected_electric_vehicle_at_charger_{id}_battery_capacity'),
    ('electric_vehicle_charger_incoming_state', 'electric_vehicle_charger_{id}_incoming_state'),
    ('incoming_electric_vehicle_at_charger_estimated_arrival_time', 'incoming_electric_vehicle_at_charger_{id}_estimated_arrival_time'),
    ('incoming_electric_vehicle_at_charger_estimated_soc_arrival', 'incoming_electric_vehicle_at_charger_{id}_estimated_soc_arrival'),
]


# per-charger parameter block (device / oracle): power limits, flat efficiency, optional charge / discharge efficiency curves
CHARGER_PARAMS = {name: i for i, name in enumerate(
    ['MAX_C', 'MIN_C', 'MAX_D', 'MIN_D', 'EFF', 'C_N', 'D_N'] + [f'C_X{j}' for j in range(8)] + [f'C_Y{j}' for j in range(8)]
    + [f'D_X{j}' for j in range(8)] + [f'D_Y{j}' for j in range(8)])}


@dataclass
class ElectricVehicleSpec:
    name: str
    battery: Dict[str, Any]                    # resolved like a building battery (schema.resolve_battery)


@dataclass
class ChargerSpec:
    charger_id: str
    building: int
    max_charging_power: float
    min_charging_power: float
    max_discharging_power: float
    min_discharging_power: float
    efficiency: float
    charge_curve: Optional[np.ndarray]         # [2, N] power level -> efficiency, or None
    discharge_curve: Optional[np.ndarray]
    # dataset series over the whole simulation period (ChargerSimulation, citylearn/data.py:699-768)
    state: np.ndarray = None                   # float: 1 connected, 2 incoming, 3 away, NaN
    ev: np.ndarray = None                      # int32 index into the district's vehicle list, -1: none / unknown id
    capacity: np.ndarray = None
    current_soc: np.ndarray = None             # raw kWh / capacity clipped to [0, 1]
    departure_time: np.ndarray = None          # int, -1 default
    required_soc: np.ndarray = None            # [0, 1] or -0.1
    arrival_time: np.ndarray = None
    soc_arrival: np.ndarray = None
    observations: Dict[str, np.ndarray] = field(default_factory=dict)     # full observation name -> float32 column


@dataclass
class WashingMachineSpec:
    name: str
    building: int
    start: np.ndarray                          # wm_start_time_step (int, -1: none)
    end: np.ndarray
    profile_sum: np.ndarray                    # float64: sum of the row's load profile (the reference adds every entry to ec[t], energy_model.py:1324-1327)
    profile_len: np.ndarray
    profiles: List[np.ndarray] = field(default_factory=list)      # the row's load profile entries
    # [n_rows][max_len - 1]: the same float32 accumulation over the first 1, 2, ... entries only - a cycle started so close to the end of
    # the episode that `t + offset` leaves it adds just those (energy_model.py:1325-1327)
    profile_prefix: np.ndarray = None

    def profile_values(self, row: int) -> np.ndarray:
        return self.profiles[row]


MAX_PHASES = 4                                 # include/citylearn_b200.h CL_MAX_PHASES


@dataclass
class ChargingConstraints:
    """`Building._initialize_charging_constraints` (citylearn/building.py:764-833): a power cap on the sum of a building's positive
    charger requests and on the chargers of each phase; the excess scales the actions down (`_apply_charging_constraints_to_actions`,
    :894-982) and is reported as `charging_constraint_violation_kwh`.  Observation names are kept in the order the reference inserts
    them into `observation_metadata`; `one_hot` maps the static phase-encoding observations to their values."""
    building_limit_kw: Optional[float]
    phases: List[Dict[str, Any]]               # {'name', 'limit_kw' (None: no cap), 'chargers': [charger ids]}
    expose_headroom: bool
    expose_violation: bool
    one_hot: Dict[str, float]
    headroom_names: List[str]                  # 'charging_building_headroom_kw', 'charging_phase_<name>_headroom_kw' (limits that exist)
    config: Dict[str, Any] = field(default_factory=dict)

    @property
    def state_names(self) -> List[str]:
        """Observations held as per-env state on the device, in slot order."""
        return (self.headroom_names if self.expose_headroom else []) + (['charging_constraint_violation_kwh'] if self.expose_violation else [])


def load_charging_constraints(config: Optional[Mapping[str, Any]], chargers: List[ChargerSpec]) -> Optional[ChargingConstraints]:
    if not config:
        return None
    oc = config.get('observations', {}) or {}
    flag = config.get('expose_observations')
    expose = bool(oc.get('headroom', False)) if 'headroom' in oc else (bool(flag) if flag is not None else True)
    phases = []
    phase_of: Dict[str, str] = {}
    for ph in config.get('phases', []) or []:
        name = ph.get('name') or f'phase_{len(phases) + 1}'
        ids = list(ph.get('chargers', []) or [])
        phases.append({'name': name, 'limit_kw': ph.get('limit_kw'), 'chargers': ids})
        for cid in ids:
            phase_of[cid] = name
    encode = bool(oc.get('phase_encoding', False)) and bool(phases)
    one_hot: Dict[str, float] = {}
    ids = [c.charger_id for c in chargers]
    if encode and ids:
        names = sorted({ph['name'] for ph in phases if ph['name']})
        unassigned = any(cid not in phase_of for cid in ids)
        if unassigned:
            names = names + ['unassigned']
        for cid in ids:
            mine = phase_of.get(cid, 'unassigned' if unassigned else None)
            for n in names:
                one_hot[f'charging_phase_one_hot_{cid}_{n}'] = 1.0 if mine == n else 0.0
    head = []
    if config.get('building_limit_kw') is not None:
        head.append('charging_building_headroom_kw')
    head += [f"charging_phase_{ph['name']}_headroom_kw" for ph in phases if ph['limit_kw'] is not None]
    return ChargingConstraints(config.get('building_limit_kw'), phases, expose, bool(oc.get('violation', True)), one_hot, head, dict(config))


def _stable_unit(name: str) -> float:
    """Stand-in for the reference's `random.uniform(0, 1)` default of a vehicle's initial SOC (unseeded there)."""
    return int(hashlib.md5(name.encode()).hexdigest()[:8], 16) / float(0xFFFFFFFF)


def load_electric_vehicles(sch: Mapping[str, Any], kwargs: Mapping[str, Any], resolve_battery) -> List[ElectricVehicleSpec]:
    """`CityLearnEnv._load_electric_vehicle` (citylearn/citylearn.py:2558-2594) for every included vehicle."""
    defs = kwargs.get('electric_vehicles_def') or sch.get('electric_vehicles_def') or {}
    out = []
    for name, es in defs.items():
        if not es.get('include'):
            continue
        a = es['battery']['attributes']
        attrs = {'capacity': a['capacity'], 'nominal_power': a['nominal_power'],
                 'initial_soc': a['initial_soc'] if a.get('initial_soc') is not None else _stable_unit(name),
                 'depth_of_discharge': a.get('depth_of_discharge', 0.10)}
        out.append(ElectricVehicleSpec(name, resolve_battery(attrs, sch.get('random_seed'))))
    return out


def _num(col, default=np.nan):
    a = np.array([default if v is None else v for v in col], dtype='float64') if not isinstance(col, np.ndarray) else col.astype('float64')
    return a


def load_chargers(building_index: int, bs: Mapping[str, Any], source, lo: int, hi: int, ev_names: List[str]) -> List[ChargerSpec]:
    """The chargers of one building (citylearn/citylearn.py:2277-2298): series sliced to the simulation period like the reference's
    `iloc[simulation_start:simulation_end + 1]`, then re-padded to dataset length so that rows index like every other series."""
    out = []
    for cid, cfg in (bs.get('chargers') or {}).items():
        if cfg.get('noise_std', 0.0):
            from .schema import UnsupportedSchemaError
            raise UnsupportedSchemaError(f"charger '{cid}': noise_std > 0 draws from NumPy's global generator at load time; not supported")
        if not cfg.get('charger_simulation'):
            raise ValueError(f"charger '{cid}': the schema names no 'charger_simulation' file")
        t = source.text_table(cfg['charger_simulation'])
        n = len(next(iter(t.values())))
        a = dict(cfg.get('attributes', {}) or {})

        def curve(key):
            v = a.get(key)
            return None if v is None else np.array(v, dtype='float64').T
        # `int(str(s)) if str(s).isdigit() else nan` on the values pandas parsed (citylearn/data.py:721-724): an all-integer column
        # arrives as ints; a column with missing cells arrives as floats, whose str() ('1.0') is never a digit string -> all NaN
        state_raw = np.asarray(t['electric_vehicle_charger_state'], dtype='float64')
        if np.isnan(state_raw).any() or (state_raw != np.floor(state_raw)).any() or (state_raw < 0).any():
            state = np.full(len(state_raw), np.nan)
        else:
            state = state_raw.copy()
        ids = t['electric_vehicle_id']
        ev = np.array([ev_names.index(str(x).strip()) if isinstance(x, str) and str(x).strip() in ev_names else -1 for x in ids], dtype='int32')
        cap = _num(t['electric_vehicle_battery_capacity_khw'])
        cur = _num(t['current_soc'])
        cur = np.where(np.isnan(cur), DEFAULT_SOC, cur)
        with np.errstate(divide='ignore', invalid='ignore'):
            current_soc = np.clip(cur / cap, 0, 1)
        dep = _num(t['electric_vehicle_departure_time'])
        dep = np.where(np.isnan(dep), DEFAULT_TIME, dep).astype('int64')
        arr = _num(t['electric_vehicle_estimated_arrival_time'])
        arr = np.where(np.isnan(arr), DEFAULT_TIME, arr).astype('int64')
        req = _num(t['electric_vehicle_required_soc_departure'])
        req = np.where(np.isnan(req), DEFAULT_SOC, req)
        req = np.where(req != DEFAULT_SOC, np.clip(req / 100 + 0.0 / 100, 0, 1), req)
        soa = _num(t['electric_vehicle_estimated_soc_arrival'])
        soa = np.where(np.isnan(soa), DEFAULT_SOC, soa)
        soa = np.where(soa != DEFAULT_SOC, np.clip(soa / 100 + 0.0 / 100, 0, 1), soa)
        c = ChargerSpec(
            charger_id=cid, building=building_index,
            max_charging_power=float(a.get('max_charging_power', 50) if a.get('max_charging_power') is not None else 50),
            min_charging_power=float(a.get('min_charging_power', 0) or 0), max_discharging_power=float(a.get('max_discharging_power', 50) if a.get('max_discharging_power') is not None else 50),
            min_discharging_power=float(a.get('min_discharging_power', 0) or 0), efficiency=float(a['efficiency']) if a.get('efficiency') is not None else float('nan'),
            charge_curve=curve('charge_efficiency_curve'), discharge_curve=curve('discharge_efficiency_curve'),
            state=state, ev=ev, capacity=cap, current_soc=current_soc, departure_time=dep, required_soc=req, arrival_time=arr, soc_arrival=soa)
        assert n == len(state)
        out.append(c)
    return out


def load_washing_machines(building_index: int, bs: Mapping[str, Any], source, kwargs) -> List[WashingMachineSpec]:
    """citylearn/citylearn.py:2300-2308, 2596-2640; `WashingMachineSimulation` citylearn/data.py:770-820."""
    out = []
    for name, cfg in (kwargs.get('washing_machines') or bs.get('washing_machines') or {}).items():
        t = source.text_table(cfg['washing_machine_energy_simulation'])
        st = _num(t['wm_start_time_step'])
        en = _num(t['wm_end_time_step'])
        st = np.where(np.isnan(st), DEFAULT_TIME, st).astype('int64')
        en = np.where(np.isnan(en), DEFAULT_TIME, en).astype('int64')
        sums, lens, profs, partial = [], [], [], []
        for s in t['load_profile']:
            try:
                p = np.array(eval(str(s), {'__builtins__': {}}, {}), dtype='float64')      # '[3.157]' -> array; '-1' -> 0-d array
                p = p.reshape(-1) if p.ndim else np.zeros(0)
            except Exception:
                p = np.zeros(0)
            acc = np.float32(0.0)
            part = []
            for x in p:                      # `ec[t] += entry`: float32 slot, np.float64 entries (energy_model.py:1324-1327)
                acc = np.float32(np.float64(acc) + np.float64(x))
                part.append(float(acc))
            sums.append(float(acc))
            partial.append(part)
            lens.append(len(p))
            profs.append(p)
        width = max(max(lens, default=1) - 1, 0)
        prefix = np.zeros((len(sums), width))
        for r, part in enumerate(partial):
            for j in range(width):           # first j + 1 entries (the full sum once the profile is exhausted)
                prefix[r, j] = part[min(j, len(part) - 1)] if part else 0.0
        out.append(WashingMachineSpec(name, building_index, st, en, np.array(sums), np.array(lens, dtype='int64'), profs, prefix))
    return out


# ------------------------------------------------------------------------------------------------------------------------------
# schedule compilation
# ------------------------------------------------------------------------------------------------------------------------------
def _valid_soc(v) -> bool:
    return isinstance(v, (float, np.floating)) and not np.isnan(v) and 0.0 <= v <= 1.0


def connected_mask(c: ChargerSpec) -> np.ndarray:
    """`charger.connected_electric_vehicle and state == 1` after `associate_chargers_to_electric_vehicles` (citylearn.py:1379-1411):
    state 1 with an id naming one of the district's vehicles."""
    return (c.state == 1) & (c.ev >= 0)


def incoming_mask(c: ChargerSpec) -> np.ndarray:
    return (c.state == 2) & (c.ev >= 0)


def arrival_soc(c: ChargerSpec, r: int, first_of_episode: bool) -> Optional[float]:
    """`_resolve_arrival_soc` (citylearn.py:1356-1377) for a vehicle plugged in at row r; None when no value is usable."""
    cand = r
    if not first_of_episode and r > 0 and c.state[r - 1] == 2 and c.ev[r - 1] == c.ev[r] and c.ev[r] >= 0:
        cand = r - 1
    v = c.soc_arrival[cand] if 0 <= cand < len(c.soc_arrival) else np.nan
    if _valid_soc(float(v)):
        return float(v)
    f = c.current_soc[min(r, len(c.current_soc) - 1)]
    return float(f) if _valid_soc(float(f)) else None


def is_new_connection(c: ChargerSpec, r: int, first_of_episode: bool) -> bool:
    if first_of_episode or r == 0:
        return True
    return not (c.state[r - 1] == 1 and c.ev[r - 1] == c.ev[r])


def compile_schedule(chargers: List[ChargerSpec], n_ev: int, n_rows: int, lo: int, hi: int, seed: Optional[int]) -> Dict[str, np.ndarray]:
    """Per (row, vehicle): what `next_time_step` does to the vehicle's `soc[t]` entry when the simulation ARRIVES at that row
    (citylearn.py:1346-1351): `assoc[r, v]` constant from `associate_chargers_to_electric_vehicles` (new connection), `sim[r, v]`
    constant from `simulate_unconnected_ev_soc` (about to connect), `drift[r, v]` factor on soc[t-1] (away); NaN = not applicable.
    `t0[r, v]`: the constant a reset at row r applies (every connection is new at t = 0).  Rows outside [lo, hi] stay NaN.

    The drift factors are drawn in (row, vehicle) order from `RandomState(seed)` - the order the reference consumes its global
    generator in a first episode that starts at row `lo`."""
    assoc = np.full((n_rows, n_ev), np.nan)
    sim = np.full((n_rows, n_ev), np.nan)
    drift = np.full((n_rows, n_ev), np.nan)
    t0 = np.full((n_rows, n_ev), np.nan)
    rs = np.random.RandomState(seed)
    for r in range(lo, hi + 1):
        # simulate_unconnected_ev_soc at t' = r (relative to an episode that started earlier): for every vehicle, first charger that matches
        if r > lo:
            for v in range(n_ev):
                found = False
                for c in chargers:
                    cur_id, cur_state = c.ev[r], c.state[r]
                    nxt_id = c.ev[r + 1] if r + 1 <= hi else -1
                    nxt_state = c.state[r + 1] if r + 1 <= hi else np.nan
                    if cur_id == v and cur_state == 1:
                        found = True
                        break
                    if nxt_id == v and nxt_state == 1 and cur_state != 1:
                        found = True
                        s = c.soc_arrival[r] if (cur_id == v and cur_state == 2) else (c.soc_arrival[r + 1] if r + 1 <= hi else np.nan)
                        if 0 <= s <= 1:
                            sim[r, v] = s
                        break
                if not found:
                    drift[r, v] = float(np.clip(rs.normal(1.0, 0.2), 0.6, 1.4))
        for c in chargers:
            if c.state[r] == 1 and c.ev[r] >= 0:
                v = int(c.ev[r])
                if is_new_connection(c, r, False):
                    s = arrival_soc(c, r, False)
                    if s is not None:
                        assoc[r, v] = s
                s0 = arrival_soc(c, r, True)
                if s0 is not None:
                    t0[r, v] = s0
    return {'assoc': assoc, 'sim': sim, 'drift': drift, 't0': t0}


def charger_observation_columns(c: ChargerSpec, sched: Mapping[str, np.ndarray], initial_soc: np.ndarray) -> Dict[str, np.ndarray]:
    """Reference-parity observation columns of one charger (citylearn/building.py:1221-1296), by schema observation name.

    `..._soc` is `battery.soc[t]` of the plugged-in vehicle at observation time, i.e. before any action at t: the arrival SOC on the
    row of a new connection, the zero-initialised entry otherwise.  (At t = 0 of an episode the env patches the reset values.)"""
    con, inc = connected_mask(c), incoming_mask(c)
    n = len(c.state)
    soc = np.full(n, DEFAULT_SOC)
    idx = np.nonzero(con)[0]
    for r in idx:
        a = sched['assoc'][r, c.ev[r]]
        soc[r] = 0.0 if np.isnan(a) else a
    f32 = lambda a: np.asarray(a, dtype='float32')  # noqa: E731
    return {
        'electric_vehicle_charger_connected_state': f32(con.astype('float64')),
        'connected_electric_vehicle_at_charger_departure_time': f32(np.where(con, c.departure_time, -1)),
        'connected_electric_vehicle_at_charger_required_soc_departure': f32(np.where(con, c.required_soc, DEFAULT_SOC)),
        'connected_electric_vehicle_at_charger_soc': f32(soc),
        'connected_electric_vehicle_at_charger_battery_capacity': f32(np.where(con, c.capacity, -1.0)),
        'electric_vehicle_charger_incoming_state': f32(inc.astype('float64')),
        'incoming_electric_vehicle_at_charger_estimated_arrival_time': f32(np.where(inc, c.arrival_time, -1)),
        'incoming_electric_vehicle_at_charger_estimated_soc_arrival': f32(np.where(inc, c.soc_arrival, DEFAULT_SOC)),
    }
