"""Synthetic wide districts (BASELINE.json configs[3], SURVEY.md §8d "C4").

Building i of an N-building district replays `Building_{(i mod 17)+1}.csv` of `citylearn_challenge_2022_phase_all` with its
non-shiftable load scaled by `s_i = RandomState(seed).uniform(0.5, 1.5, N)[i]`, a 6.4 kWh / 5 kW battery (efficiency 0.9,
capacity loss 1e-5, no self-discharge; curves are the reference's md5-seeded defaults, so they differ per building name) and
a `4 + (i mod 2)` kW PV array.  Weather, pricing and carbon intensity are shared.

`SyntheticWideSource` serves the district from memory; `write_directory` writes the same thing as a real schema directory
(`schema.json` + CSV files) so that the unmodified reference can be run on it (oracle/make_golden.py uses a 32-building one).
"""
from __future__ import annotations

import copy
import json
import os
from pathlib import Path
from typing import Dict, Optional

import numpy as np

from .data import DataSet, DataSource

BASE = 'citylearn_challenge_2022_phase_all'
N_BASE = 17


class SyntheticWideSource(DataSource):
    def __init__(self, n_buildings: int, seed: int = 0, base: Optional[DataSource] = None):
        assert n_buildings >= 1
        self.n = int(n_buildings)
        self.base = DataSet.get_source(BASE) if base is None else base
        self.scale = np.random.RandomState(seed).uniform(0.5, 1.5, self.n)

    @staticmethod
    def file_of(i: int) -> str:
        return f'Synthetic_{i + 1}.csv'

    def root_directory(self):
        return None

    def schema(self) -> dict:
        s = copy.deepcopy(self.base.schema())
        proto = s['buildings']['Building_1']
        buildings = {}
        for i in range(self.n):
            b = copy.deepcopy(proto)
            b['energy_simulation'] = self.file_of(i)
            b['electrical_storage']['attributes'] = {'capacity': 6.4, 'efficiency': 0.9, 'capacity_loss_coefficient': 1e-05,
                                                    'loss_coefficient': 0.0, 'nominal_power': 5.0}
            b['pv']['attributes'] = {'nominal_power': 4.0 + (i % 2)}
            buildings[f'Building_{i + 1}'] = b
        s['buildings'] = buildings
        s['root_directory'] = None
        return s

    def table(self, filename: str) -> Dict[str, np.ndarray]:
        if filename.startswith('Synthetic_'):
            i = int(filename[len('Synthetic_'):-len('.csv')]) - 1
            t = dict(self.base.table(f'Building_{(i % N_BASE) + 1}.csv'))
            # what a CSV round trip of the scaled float32 series gives: float64 text -> float32 in the loader
            t['non_shiftable_load'] = np.asarray(t['non_shiftable_load'], dtype='float64') * self.scale[i]
            return t
        return self.base.table(filename)

    def state_dict(self, filename: str):
        return self.base.state_dict(filename)

    def write_directory(self, root: os.PathLike) -> Path:
        """Materialise the district as `root/schema.json` + CSV files (lossless decimal text)."""
        return _write_directory(self, root)


def _schema_files(sch: dict):
    files = set()
    for b in sch['buildings'].values():
        for k in ('energy_simulation', 'weather', 'carbon_intensity', 'pricing'):
            if b.get(k):
                files.add(b[k])
    return files


def _write_directory(src: DataSource, root: os.PathLike) -> Path:
    root = Path(root)
    root.mkdir(parents=True, exist_ok=True)
    sch = src.schema()
    files = _schema_files(sch)
    for fn in sorted(files):
        t = src.table(fn)
        cols = list(t)
        arr = np.stack([np.asarray(t[c], dtype='float64') for c in cols], axis=1)
        with open(root / fn, 'w') as f:
            f.write(','.join(cols) + '\n')
            for r in arr:
                f.write(','.join('' if np.isnan(v) else repr(float(v)) for v in r) + '\n')
    for b in sch['buildings'].values():          # LSTM dynamics weights (`.pth` state dicts, citylearn/dynamics.py:112-127)
        fn = ((b.get('dynamics') or {}).get('attributes') or {}).get('filename')
        if fn:
            import torch
            torch.save({k: torch.as_tensor(np.asarray(v)) for k, v in src.state_dict(fn).items()}, root / fn)
    sch['root_directory'] = None
    with open(root / 'schema.json', 'w') as f:
        json.dump(sch, f, indent=1)
    return root


def make_wide_district(n_buildings: int = 1024, seed: int = 0):
    """(schema dict, data source) of the synthetic N-building district; pass both to `CityLearnEnv(schema, data_source=...)`."""
    src = SyntheticWideSource(n_buildings, seed)
    return src.schema(), src


class SyntheticHeatingSource(DataSource):
    """`citylearn_challenge_2020_climate_zone_1` with a heating season: no bundled dataset has space-heating demand, so the
    heating heat pump, the heating tank and the reference's tank-capacity quirks (heating storage actions scale with the COOLING
    tank, DHW storage actions with the HEATING tank, `building.py:1720, 1765`) would otherwise never run.  Every building serves
    0.6 x its December-February cooling load as HEATING demand instead, and gets an autosized heating heat pump and an autosized heating tank;
    the `heating_storage` action and `heating_storage_soc` observation are switched on.  Test / parity infrastructure, like
    `SyntheticWideSource`."""

    def __init__(self, base: Optional[DataSource] = None):
        self.base = DataSet.get_source('citylearn_challenge_2020_climate_zone_1') if base is None else base

    def root_directory(self):
        return None

    def schema(self) -> dict:
        s = copy.deepcopy(self.base.schema())
        s['observations']['heating_storage_soc']['active'] = True
        s['actions']['heating_storage']['active'] = True
        for b in s['buildings'].values():
            b['heating_device'] = {'type': 'citylearn.energy_model.HeatPump', 'autosize': True,
                                   'attributes': {'nominal_power': None, 'efficiency': 0.25, 'target_cooling_temperature': 8.0,
                                                  'target_heating_temperature': 45.0}}
            b['heating_storage'] = {'type': 'citylearn.energy_model.StorageTank', 'autosize': True, 'autosize_attributes': {'safety_factor': 1.5},
                                    'attributes': {'capacity': None, 'loss_coefficient': 0.004}}
        s['root_directory'] = None
        return s

    def table(self, filename: str) -> Dict[str, np.ndarray]:
        t = dict(self.base.table(filename))
        if 'cooling_demand' in t and 'heating_demand' in t:
            # December-February become a heating season (the reference forbids cooling and heating in the same time step,
            # citylearn/data.py:470-472): the cooling load of those months, scaled, is served as heat instead
            cool = np.asarray(t['cooling_demand'], dtype='float32')
            winter = np.isin(np.asarray(t['month']).astype(int), (12, 1, 2))
            t['heating_demand'] = np.where(winter, np.float32(0.6) * cool, np.float32(0.0)).astype('float64')
            t['cooling_demand'] = np.where(winter, np.float32(0.0), cool).astype('float64')
        return t

    def state_dict(self, filename: str):
        return self.base.state_dict(filename)

    def write_directory(self, root: os.PathLike) -> Path:
        return _write_directory(self, root)


class SyntheticDualModeSource(DataSource):
    """`baeda_3dem` (buildings 1-3, LSTM dynamics) driven through ONE signed `cooling_or_heating_device` action
    (`building.py:1550-1553`: negative = cooling, positive = heating share of the nominal power) instead of `cooling_device`.
    The night hours (0-5 h) in which the HVAC is off become heating hours (`hvac_mode = 2`, a synthetic heating demand of 0.4 x the
    building's mean cooling load) served by an autosized heating heat pump, and every 7th cooling hour runs in auto mode
    (`hvac_mode = 3`).  The only reference datasets using this action need PV autosizing through PySAM.  Parity infrastructure."""

    BUILDINGS = ['Building_1', 'Building_2', 'Building_3']

    def __init__(self, base: Optional[DataSource] = None):
        self.base = DataSet.get_source('baeda_3dem') if base is None else base

    def root_directory(self):
        return None

    def schema(self) -> dict:
        s = copy.deepcopy(self.base.schema())
        s['buildings'] = {k: v for k, v in s['buildings'].items() if k in self.BUILDINGS}
        s['actions']['cooling_device']['active'] = False
        s['actions']['cooling_or_heating_device']['active'] = True
        for b in s['buildings'].values():
            # oversized: the reference books the ideal load at t = 0 twice, which eats into the heat pump's head-room
            b['heating_device'] = {'type': 'citylearn.energy_model.HeatPump', 'autosize': True, 'autosize_attributes': {'safety_factor': 3.0},
                                   'attributes': {'nominal_power': None, 'efficiency': 0.25, 'target_cooling_temperature': 8.0,
                                                  'target_heating_temperature': 45.0}}
            b['inactive_actions'] = [a for a in b.get('inactive_actions', []) if a != 'cooling_or_heating_device']
        s['root_directory'] = None
        return s

    def table(self, filename: str) -> Dict[str, np.ndarray]:
        t = dict(self.base.table(filename))
        if 'hvac_mode' in t and 'cooling_demand' in t:
            cool = np.nan_to_num(np.asarray(t['cooling_demand'], dtype='float32'))
            mode = np.nan_to_num(np.asarray(t['hvac_mode'])).astype(int)
            hour = np.asarray(t['hour']).astype(int)
            heat_rows = (mode == 0) & (hour <= 5) & (cool == 0)
            level = np.float32(0.4) * np.float32(cool[cool > 0].mean())
            t['heating_demand'] = np.where(heat_rows, level, np.float32(0.0)).astype('float64')
            auto_rows = np.zeros(len(mode), dtype=bool)
            auto_rows[np.nonzero(mode == 1)[0][::7]] = True
            t['hvac_mode'] = np.where(heat_rows, 2, np.where(auto_rows, 3, mode)).astype('float64')
        return t

    def state_dict(self, filename: str):
        return self.base.state_dict(filename)

    def write_directory(self, root: os.PathLike) -> Path:
        return _write_directory(self, root)
