"""Dataset access for the B200 CityLearn hot path.

The reference resolves a dataset *name* by downloading it from GitHub
(`citylearn/data.py:113-189`) and a schema *path* by reading `schema.json` plus the
CSV / `.pth` files next to it (`citylearn/citylearn.py:2183-2207`, `citylearn/dynamics.py:112-127`).
There is no network on a B200 box, so names resolve to compact `.npz` packs bundled with
this package (`citylearn_b200/datasets/<name>.npz`, produced by `tools/pack_datasets.py`
from the same CSV files), while paths are read exactly like the reference does.

Every series is ingested the way the reference ingests it: `pandas.read_csv` -> float64 ->
`np.array(..., dtype='float32')` (`citylearn/data.py:399-493,515-595,599-661`), so a pack
stores float32 and is lossless with respect to what the reference computes with.
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

__all__ = ['DataSource', 'DirectorySource', 'PackSource', 'DataSet', 'TOLERANCE', 'ZERO_DIVISION_PLACEHOLDER']

# citylearn/data.py:18-19
TOLERANCE = 0.0001
ZERO_DIVISION_PLACEHOLDER = 0.000001

_PACK_DIR = Path(__file__).resolve().parent / 'datasets'


class DataSource:
    """Where the files a schema refers to come from."""

    def schema(self) -> dict:
        raise NotImplementedError

    def table(self, filename: str) -> Dict[str, np.ndarray]:
        """Columns of a CSV file as float64 arrays (NaN for empty cells), in file order."""
        raise NotImplementedError

    def state_dict(self, filename: str) -> Dict[str, np.ndarray]:
        """`model_state_dict` of an LSTM dynamics file as float32 arrays."""
        raise NotImplementedError

    def text_table(self, filename: str) -> Dict[str, object]:
        """Columns of a CSV file that may hold text (charger schedules name vehicles, washing machines list load profiles): numeric
        columns as float64 arrays (NaN for empty cells), the others as lists of `str` / None."""
        raise NotImplementedError

    @property
    def root_directory(self) -> Optional[str]:
        return None


class DirectorySource(DataSource):
    """schema.json + CSV + .pth in one directory (the reference's on-disk layout)."""

    def __init__(self, root: os.PathLike, schema: Optional[dict] = None):
        self._root = str(root)
        self._schema = schema
        self._cache: Dict[str, Dict[str, np.ndarray]] = {}

    @property
    def root_directory(self):
        return self._root

    def schema(self) -> dict:
        if self._schema is None:
            with open(os.path.join(self._root, 'schema.json')) as f:
                self._schema = json.load(f)
        return self._schema

    def table(self, filename: str) -> Dict[str, np.ndarray]:
        if filename not in self._cache:
            import pandas as pd
            df = pd.read_csv(os.path.join(self._root, filename))
            self._cache[filename] = {c: df[c].to_numpy(dtype='float64', na_value=np.nan) for c in df.columns}
        return self._cache[filename]

    def text_table(self, filename: str) -> Dict[str, object]:
        import pandas as pd
        df = pd.read_csv(os.path.join(self._root, filename))
        out: Dict[str, object] = {}
        for c in df.columns:
            col = df[c]
            if pd.api.types.is_numeric_dtype(col):
                out[c] = col.to_numpy(dtype='float64', na_value=np.nan)
            else:
                out[c] = [None if (v is None or (isinstance(v, float) and np.isnan(v)) or v is pd.NA) else str(v) for v in col.tolist()]
        return out

    def state_dict(self, filename: str) -> Dict[str, np.ndarray]:
        import torch
        path = os.path.join(self._root, filename)
        obj = torch.load(path, map_location='cpu', weights_only=False)
        sd = obj['model_state_dict'] if isinstance(obj, dict) and 'model_state_dict' in obj else obj
        return {k: v.detach().cpu().numpy().astype('float32') for k, v in sd.items()}


class PackSource(DataSource):
    """A dataset packed into one `.npz` (see `tools/pack_datasets.py`)."""

    def __init__(self, path: os.PathLike):
        self._path = str(path)
        self._npz = np.load(self._path, allow_pickle=False)
        self._index = json.loads(bytes(self._npz['__index__']).decode())
        self._tables: Dict[str, Dict[str, np.ndarray]] = {}

    def schema(self) -> dict:
        return json.loads(bytes(self._npz['__schema__']).decode())

    def table(self, filename: str) -> Dict[str, np.ndarray]:
        if filename not in self._tables:        # decompress once; callers get a fresh dict of read-only arrays
            cols = self._index['tables'][filename]
            t = {c: self._npz[f'{filename}::{c}'].astype('float64') for c in cols}
            for a in t.values():
                a.setflags(write=False)
            self._tables[filename] = t
        return dict(self._tables[filename])

    def text_table(self, filename: str) -> Dict[str, object]:
        cols = self._index['text_tables'][filename]
        out: Dict[str, object] = {}
        for c, kind in cols.items():
            if kind == 'num':
                out[c] = self._npz[f'{filename}::{c}'].astype('float64')
            else:
                out[c] = json.loads(bytes(self._npz[f'{filename}::{c}']).decode())
        return out

    def state_dict(self, filename: str) -> Dict[str, np.ndarray]:
        keys = self._index['state_dicts'][filename]
        return {k: self._npz[f'{filename}::{k}'] for k in keys}


def write_pack(src: DirectorySource, out_path: os.PathLike) -> None:
    """Pack every file the schema of `src` refers to into one compressed `.npz`."""
    schema = src.schema()
    arrays: Dict[str, np.ndarray] = {}
    index = {'tables': {}, 'state_dicts': {}, 'text_tables': {}}

    def add_table(fn):
        if fn is None or fn in index['tables']:
            return
        t = src.table(fn)
        index['tables'][fn] = list(t.keys())
        for c, v in t.items():
            arrays[f'{fn}::{c}'] = v.astype('float32')

    def add_text_table(fn):
        if fn is None or fn in index['text_tables']:
            return
        t = src.text_table(fn)
        index['text_tables'][fn] = {}
        for c, v in t.items():
            if isinstance(v, np.ndarray):
                index['text_tables'][fn][c] = 'num'
                arrays[f'{fn}::{c}'] = v.astype('float64')      # schedules are small; keep the parsed values exactly
            else:
                index['text_tables'][fn][c] = 'text'
                arrays[f'{fn}::{c}'] = np.frombuffer(json.dumps(v).encode(), dtype='uint8')

    for b in schema['buildings'].values():
        for k in ('energy_simulation', 'weather', 'carbon_intensity', 'pricing'):
            add_table(b.get(k))
        for cfg in (b.get('chargers') or {}).values():
            add_text_table(cfg.get('charger_simulation'))
        for cfg in (b.get('washing_machines') or {}).values():
            add_text_table(cfg.get('washing_machine_energy_simulation'))
        dyn = b.get('dynamics')
        if dyn is not None:
            fn = dyn['attributes']['filename']
            if fn not in index['state_dicts']:
                sd = src.state_dict(fn)
                index['state_dicts'][fn] = list(sd.keys())
                for k, v in sd.items():
                    arrays[f'{fn}::{k}'] = v
    arrays['__schema__'] = np.frombuffer(json.dumps(schema).encode(), dtype='uint8')
    arrays['__index__'] = np.frombuffer(json.dumps(index).encode(), dtype='uint8')
    np.savez_compressed(out_path, **arrays)


class DataSet:
    """Offline stand-in for `citylearn.data.DataSet` (`citylearn/data.py:31-292`): names map to bundled packs."""

    @staticmethod
    def get_dataset_names() -> List[str]:
        return sorted(p.stem for p in _PACK_DIR.glob('*.npz'))

    @staticmethod
    def get_source(name: str) -> PackSource:
        path = _PACK_DIR / f'{name}.npz'
        if not path.is_file():
            raise FileNotFoundError(
                f"dataset '{name}' is not bundled (bundled: {DataSet.get_dataset_names()}); "
                'pass the path of a schema.json instead (no network access to download datasets).')
        return PackSource(path)

    @staticmethod
    def get_schema(name: str) -> dict:
        return DataSet.get_source(name).schema()
