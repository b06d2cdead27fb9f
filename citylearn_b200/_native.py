"""ctypes binding of `libcitylearn_b200.so` (C ABI in `include/citylearn_b200.h`).

The CUDA library is the only compute path of this package: if it is missing or cannot be loaded this
module raises - there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import Optional

import numpy as np

from . import schema as S

_LIB_NAME = 'libcitylearn_b200.so'
_lib = None


class NativeLibraryError(RuntimeError):
    pass


class DistrictDesc(ctypes.Structure):
    _fields_ = [
        ('abi_version', ctypes.c_int32), ('n_buildings', ctypes.c_int32), ('n_envs', ctypes.c_int32),
        ('n_rows', ctypes.c_int32), ('n_cols', ctypes.c_int32), ('action_dim', ctypes.c_int32),
        ('obs_dim', ctypes.c_int32), ('central_agent', ctypes.c_int32), ('reward_id', ctypes.c_int32),
        ('precision', ctypes.c_int32), ('stale_observations', ctypes.c_int32), ('lstm_weight_count', ctypes.c_int32),
        ('reward_params', ctypes.c_double * 8),
        ('table', ctypes.c_void_p), ('params', ctypes.c_void_p), ('iparams', ctypes.c_void_p),
        ('obs_desc', ctypes.c_void_p), ('lstm_weights', ctypes.c_void_p), ('ev', ctypes.c_void_p),
    ]


class EvDesc(ctypes.Structure):          # cl_ev_desc
    _fields_ = [('n_ev', ctypes.c_int32), ('n_chargers', ctypes.c_int32), ('n_machines', ctypes.c_int32)] + [
        (n, ctypes.c_void_p) for n in ('ev_params', 'ev_iparams', 'ev_cols', 'ev_drift', 'ch_building', 'ch_action', 'ch_cols', 'ch_params',
                                       'wm_building', 'wm_action', 'wm_cols')] + [('n_constrained', ctypes.c_int32)] + [
        (n, ctypes.c_void_p) for n in ('cc_building', 'cc_limits', 'cc_members', 'cc_flags')]


ABI_VERSION = 2
PRECISION = {'fp32': 0, 'fp64': 1}


def library_path() -> Path:
    override = os.environ.get('CL_B200_LIB')      # debug builds (e.g. tools/phase_timing.py)
    return Path(override) if override else Path(__file__).resolve().parent / _LIB_NAME


def load():
    """Load the CUDA library once; raises `NativeLibraryError` when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.is_file():
        raise NativeLibraryError(
            f'{path} not found: the CUDA extension has not been built. Run `python -c "import __graft_entry__ as g; g.build()"` '
            f'(or `make -C {path.parent / "csrc"}`) - citylearn_b200 has no CPU fallback.')
    try:
        lib = ctypes.CDLL(str(path))
    except OSError as e:  # pragma: no cover
        raise NativeLibraryError(f'cannot load {path}: {e}') from e
    vp, i32, i64p = ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)
    lib.cl_abi_version.restype = ctypes.c_int
    lib.cl_last_error.restype = ctypes.c_char_p
    lib.cl_create.argtypes = [ctypes.POINTER(DistrictDesc), ctypes.POINTER(vp)]
    lib.cl_destroy.argtypes = [vp]
    lib.cl_set_outage.argtypes = [vp, vp, i32, vp]
    lib.cl_reset.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.cl_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.cl_rollout.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.cl_device_time_enable.argtypes = [vp, vp]
    lib.cl_ev_read.argtypes = [vp, vp, vp, vp]
    lib.cl_step_host.argtypes = [vp] * 9 + [ctypes.c_size_t, i32, vp]
    lib.cl_exchange_create.argtypes = [vp, i32, i32, vp, ctypes.POINTER(vp)]
    lib.cl_exchange_connect.argtypes = [vp, vp]
    lib.cl_exchange_connect_ptrs.argtypes = [vp, ctypes.POINTER(vp), vp]
    lib.cl_exchange_status.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint32)]
    lib.cl_advance_device.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.cl_obs_rows.argtypes = [vp, i32, i32, vp, vp]
    lib.cl_time_step.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
    lib.cl_state_size.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t)]
    lib.cl_get_state.argtypes = [vp, vp, vp]
    lib.cl_set_state.argtypes = [vp, vp, i32, vp]
    lib.cl_launch_count.argtypes = [vp, i64p]
    i32p = ctypes.POINTER(ctypes.c_int32)
    lib.cl_launch_geometry.argtypes = [vp, i32p, i32p, i32p]
    lib.cl_launch_occupancy.argtypes = [vp, i32p, i32p]
    lib.cl_set_transforms.argtypes = [vp, vp, vp, vp]
    lib.cl_kpi_enable.argtypes = [vp, i32]
    lib.cl_kpi_accumulate.argtypes = [vp, vp, vp, vp]
    lib.cl_kpi_read.argtypes = [vp, vp, vp, vp]
    lib.cl_kpi_fused.argtypes = [vp, i32p]
    lib.cl_measure_fma_peak.argtypes = [ctypes.POINTER(ctypes.c_double)]
    for name in ('cl_create', 'cl_destroy', 'cl_set_outage', 'cl_reset', 'cl_step', 'cl_rollout', 'cl_obs_rows', 'cl_time_step',
                 'cl_state_size', 'cl_get_state', 'cl_set_state', 'cl_launch_count', 'cl_launch_geometry', 'cl_set_transforms',
                 'cl_kpi_enable', 'cl_kpi_accumulate', 'cl_kpi_read', 'cl_measure_fma_peak', 'cl_device_time_enable', 'cl_advance_device',
                 'cl_launch_occupancy', 'cl_kpi_fused', 'cl_ev_read', 'cl_step_host', 'cl_exchange_create', 'cl_exchange_connect', 'cl_exchange_connect_ptrs', 'cl_exchange_status'):
        getattr(lib, name).restype = ctypes.c_int
    if lib.cl_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f'{path}: ABI version {lib.cl_abi_version()} != {ABI_VERSION}; rebuild the extension')
    _lib = lib
    return lib


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = load().cl_last_error().decode(errors='replace')
        exc = {1: ValueError, 3: NotImplementedError}.get(rc, RuntimeError)
        raise exc(f'{what}: {msg}' if what else msg)


def measure_fma_peak() -> float:
    """Measured FP32 FMA throughput of the current CUDA device, TFLOP/s."""
    v = ctypes.c_double()
    check(load().cl_measure_fma_peak(ctypes.byref(v)), 'cl_measure_fma_peak')
    return v.value


class Handle:
    """Owns one `cl_env*`."""

    def __init__(self, spec: S.DistrictSpec, num_envs: int, obs_desc: np.ndarray, central_agent: bool, reward_id: int,
                 reward_params, precision: str = 'fp64', stale_observations: bool = True):
        self.lib = load()
        # keep the host arrays alive for the duration of cl_create
        table = np.ascontiguousarray(spec.table, dtype='float32')
        params = np.ascontiguousarray(spec.params.T, dtype='float64')      # [NPARAM][B]
        iparams = np.ascontiguousarray(spec.iparams.T, dtype='int32')      # [NIPARAM][B]
        desc_arr = np.ascontiguousarray(obs_desc, dtype='int32')
        weights = np.ascontiguousarray(spec.lstm_weights, dtype='float32')
        d = DistrictDesc()
        d.abi_version = ABI_VERSION
        d.n_buildings = spec.n_buildings
        d.n_envs = int(num_envs)
        d.n_rows, d.n_cols = table.shape
        d.action_dim = spec.action_dim
        d.obs_dim = desc_arr.shape[0]
        d.central_agent = int(bool(central_agent))
        d.reward_id = int(reward_id)
        d.precision = PRECISION[precision]
        d.stale_observations = int(bool(stale_observations))
        d.lstm_weight_count = int(weights.size)
        rp = list(reward_params) + [0.0] * (8 - len(reward_params))
        for i in range(8):
            d.reward_params[i] = float(rp[i])
        d.table = table.ctypes.data
        d.params = params.ctypes.data
        d.iparams = iparams.ctypes.data
        d.obs_desc = desc_arr.ctypes.data
        d.lstm_weights = weights.ctypes.data if weights.size else None
        evd = getattr(spec, 'ev', None)
        keep = []
        if evd and (len(evd['chargers']) or len(evd['wms'])):
            e = EvDesc()
            e.n_ev, e.n_chargers, e.n_machines = int(evd['n_ev']), len(evd['chargers']), len(evd['wms'])
            drift = np.ascontiguousarray(evd['schedule']['drift'], dtype='float64') if evd['n_ev'] else np.zeros((table.shape[0], 1))
            arrays = {'ev_params': (evd['ev_params'], 'float64'), 'ev_iparams': (evd['ev_ip'], 'int32'), 'ev_cols': (evd['ev_cols'], 'int32'),
                      'ev_drift': (drift, 'float64'), 'ch_building': (evd['ch_building'], 'int32'), 'ch_action': (evd['ch_action'], 'int32'),
                      'ch_cols': (evd['ch_cols'], 'int32'), 'ch_params': (evd['ch_params'], 'float64'), 'wm_building': (evd['wm_building'], 'int32'),
                      'wm_action': (evd['wm_action'], 'int32'), 'wm_cols': (evd['wm_cols'], 'int32')}
            ccb = evd.get('cc_building', ())
            e.n_constrained = len(ccb)
            if len(ccb):
                flags = [1 if spec.buildings[int(b)].charging_constraints.expose_violation else 0 for b in ccb]
                arrays.update({'cc_building': (ccb, 'int32'), 'cc_limits': (evd['cc_limits'], 'float64'), 'cc_members': (evd['cc_members'], 'int32'),
                               'cc_flags': (flags, 'int32')})
            for name, (arr, dt) in arrays.items():
                a = np.ascontiguousarray(arr, dtype=dt)
                keep.append(a)
                setattr(e, name, a.ctypes.data if a.size else None)
            keep.append(e)
            d.ev = ctypes.addressof(e)
        else:
            d.ev = None
        out = ctypes.c_void_p()
        check(self.lib.cl_create(ctypes.byref(d), ctypes.byref(out)), 'cl_create')
        self.ptr = out

    def close(self):
        if getattr(self, 'ptr', None):
            self.lib.cl_destroy(self.ptr)
            self.ptr = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_outage(self, signals, stream: int):
        if signals is None:
            check(self.lib.cl_set_outage(self.ptr, None, 0, stream), 'cl_set_outage')
        else:
            s = np.ascontiguousarray(signals, dtype='float32')
            check(self.lib.cl_set_outage(self.ptr, s.ctypes.data, s.shape[1], stream), 'cl_set_outage')

    def reset(self, start_ptr, uniform_start: int, episode_time_steps: int, obs_ptr, stream: int):
        check(self.lib.cl_reset(self.ptr, start_ptr, int(uniform_start), int(episode_time_steps), obs_ptr, stream), 'cl_reset')

    def step(self, actions_ptr, obs_ptr, reward_ptr, district_ptr, trace_ptr, stream: int):
        check(self.lib.cl_step(self.ptr, actions_ptr, obs_ptr, reward_ptr, district_ptr, trace_ptr, stream), 'cl_step')

    def rollout(self, n_steps: int, actions_ptr, obs_ptr, reward_ptr, district_ptr, stream: int):
        check(self.lib.cl_rollout(self.ptr, int(n_steps), actions_ptr, obs_ptr, reward_ptr, district_ptr, stream), 'cl_rollout')

    def step_host(self, actions_host_ptr, actions_dev_ptr, obs_ptr, reward_ptr, district_ptr, row_ptr, d2h_src_ptr, d2h_dst_ptr, d2h_bytes: int,
                  in_place: int, stream: int):
        rc = self.lib.cl_step_host(self.ptr, actions_host_ptr, actions_dev_ptr, obs_ptr, reward_ptr, district_ptr, row_ptr, d2h_src_ptr, d2h_dst_ptr,
                                   d2h_bytes, in_place, stream)
        if rc:
            check(rc, 'cl_step_host')

    def ev_read(self, soc_prev_ptr, soc_ptr, stream: int):
        check(self.lib.cl_ev_read(self.ptr, soc_prev_ptr, soc_ptr, stream), 'cl_ev_read')

    def device_time_enable(self, stream: int):
        check(self.lib.cl_device_time_enable(self.ptr, stream), 'cl_device_time_enable')

    def advance_device(self, n_steps: int, actions_ptr, obs_ptr, reward_ptr, district_ptr, stream: int):
        """Capturable (CUDA graph) variant of step / rollout: the time step lives on the device."""
        check(self.lib.cl_advance_device(self.ptr, int(n_steps), actions_ptr, obs_ptr, reward_ptr, district_ptr, stream), 'cl_advance_device')

    def exchange_create(self, n_ranks: int, rank: int):
        """-> (ipc handle bytes [64], own buffer device pointer)."""
        h = (ctypes.c_ubyte * 64)()
        buf = ctypes.c_void_p()
        check(self.lib.cl_exchange_create(self.ptr, int(n_ranks), int(rank), h, ctypes.byref(buf)), 'cl_exchange_create')
        return bytes(h), buf.value

    def exchange_connect(self, handles: bytes):
        check(self.lib.cl_exchange_connect(self.ptr, ctypes.c_char_p(handles)), 'cl_exchange_connect')

    def exchange_connect_ptrs(self, buffers, devices):
        n = len(buffers)
        arr = (ctypes.c_void_p * n)(*buffers)
        dv = np.ascontiguousarray(devices, dtype='int32')
        check(self.lib.cl_exchange_connect_ptrs(self.ptr, arr, dv.ctypes.data), 'cl_exchange_connect_ptrs')

    def exchange_status(self):
        t, e = ctypes.c_int32(), ctypes.c_uint32()
        check(self.lib.cl_exchange_status(self.ptr, ctypes.byref(t), ctypes.byref(e)), 'cl_exchange_status')
        return {'timeouts': t.value, 'steps_exchanged': e.value}

    def obs_rows(self, first_time_step: int, n_rows: int, rows_ptr, stream: int):
        check(self.lib.cl_obs_rows(self.ptr, int(first_time_step), int(n_rows), rows_ptr, stream), 'cl_obs_rows')

    def time_step(self) -> int:
        t = ctypes.c_int32()
        check(self.lib.cl_time_step(self.ptr, ctypes.byref(t)))
        return t.value

    def state_size(self) -> int:
        n = ctypes.c_size_t()
        check(self.lib.cl_state_size(self.ptr, ctypes.byref(n)))
        return n.value

    def get_state(self, dst_ptr, stream: int):
        check(self.lib.cl_get_state(self.ptr, dst_ptr, stream), 'cl_get_state')

    def set_state(self, src_ptr, time_step: int, stream: int):
        check(self.lib.cl_set_state(self.ptr, src_ptr, int(time_step), stream), 'cl_set_state')

    def set_transforms(self, obs_transform: Optional[np.ndarray], action_range: Optional[np.ndarray], action_low: Optional[np.ndarray]):
        """Fused wrapper semantics (host arrays, copied by the call): cl_obs_transform records [L]; action range / low [A]."""
        t = None if obs_transform is None else np.ascontiguousarray(obs_transform, dtype=S.OBS_TRANSFORM_DTYPE)
        r = None if action_range is None else np.ascontiguousarray(action_range, dtype='float32')
        l = None if action_low is None else np.ascontiguousarray(action_low, dtype='float32')
        check(self.lib.cl_set_transforms(self.ptr, None if t is None else t.ctypes.data, None if r is None else r.ctypes.data,
                                         None if l is None else l.ctypes.data), 'cl_set_transforms')

    def kpi_enable(self, enable: bool = True):
        check(self.lib.cl_kpi_enable(self.ptr, int(bool(enable))), 'cl_kpi_enable')

    def kpi_fused(self) -> bool:
        f = ctypes.c_int32()
        check(self.lib.cl_kpi_fused(self.ptr, ctypes.byref(f)), 'cl_kpi_fused')
        return bool(f.value)

    def kpi_accumulate(self, trace_ptr, district_ptr, stream: int):
        check(self.lib.cl_kpi_accumulate(self.ptr, trace_ptr, district_ptr, stream), 'cl_kpi_accumulate')

    def kpi_read(self, unit_ptr, env_ptr, stream: int):
        check(self.lib.cl_kpi_read(self.ptr, unit_ptr, env_ptr, stream), 'cl_kpi_read')

    def geometry(self):
        b, t, n = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        check(self.lib.cl_launch_geometry(self.ptr, ctypes.byref(b), ctypes.byref(t), ctypes.byref(n)))
        out = {'blocks': b.value, 'threads': t.value, 'tiles': n.value}
        o, sm = ctypes.c_int32(), ctypes.c_int32()
        if self.lib.cl_launch_occupancy(self.ptr, ctypes.byref(o), ctypes.byref(sm)) == 0:
            out.update(blocks_per_sm=o.value, smem_bytes=sm.value)
        return out

    @property
    def tiles(self) -> int:
        return self.geometry()['tiles']

    def launch_count(self) -> int:
        n = ctypes.c_int64()
        check(self.lib.cl_launch_count(self.ptr, ctypes.byref(n)))
        return n.value
