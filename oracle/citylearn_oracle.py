"""CPU oracle: a NumPy restatement of the reference's per-timestep simulation + reward path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this module, and only as the checker / the CPU baseline - never
as a product path (the product fails loudly when the CUDA library is missing).

Pinned (parity is NOT unpinned): `tests/test_oracle_golden.py` checks this restatement against traces
produced by running the unmodified reference itself (`oracle/make_golden.py` -> `tests/golden/*.npz`).

Arithmetic: float64 intermediates with float32 rounding wherever the reference stores into its float32
per-device arrays (soc, energy_balance, electricity_consumption, net/cost/emission ...).  That is the
reference's own precision when it is fed Python-float actions under NumPy 2, so the
oracle tracks the reference to ~1e-7 and measures how far the fp32 CUDA path is from it.

Vectorised over units `[E, B]` (E parallel envs x B buildings).  Follows, in order:
  CityLearnEnv.step .............. citylearn/citylearn.py:978-1056
  Building.apply_actions ......... citylearn/building.py:1500-1634 (priority list, :1606-1622)
  update_energy_from_*_device .... citylearn/building.py:1641,1694,1739
  update_*_storage ............... citylearn/building.py:1663,1711,1756 (wrong-tank capacities kept, :1720,:1765)
  update_non_shiftable_load ...... citylearn/building.py:1784
  update_electrical_storage ...... citylearn/building.py:1791-1812
  downward_electrical_flexibility  citylearn/building.py:639-668
  StorageDevice/StorageTank/Battery citylearn/energy_model.py:603-870, 872-1242
  HeatPump / ElectricHeater ...... citylearn/energy_model.py:216-307, 378-423
  LSTMDynamicsBuilding ........... citylearn/building.py:2935-3158, citylearn/dynamics.py:94-127
  Building.update_variables ...... citylearn/building.py:2615-2703 (t == 0 multi-counting kept)
  district sums .................. citylearn/citylearn.py:1888-1918
  reward functions ............... citylearn/reward_function.py:65-386
  observations ................... citylearn/building.py:1336-1481, citylearn/citylearn.py:451-485
"""
from __future__ import annotations

import numpy as np

from citylearn_b200 import schema as S
from citylearn_b200.schema import DYN, EPS, IP, NDYN, P

f64 = np.float64


def r32(x):
    """Round to float32 (a store into one of the reference's float32 arrays), keep computing in float64."""
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def w32(x):
    """A Python float meeting a float32 operand is converted to float32 first (NumPy 2 weak-scalar promotion)."""
    return np.asarray(x, dtype=np.float32)


def _interp_curve(xn, xs, ys, n):
    """`Battery.get_current_efficiency` / `get_max_input_power` lookup (citylearn/energy_model.py:1083-1109).

    idx = max(0, argmax(xn <= xs) - 1): argmax of an all-False mask is 0 -> segment 0 is extrapolated.
    xs, ys: [B, MAX_CURVE]; n: [B] number of valid points; xn: [E, B].
    """
    E, B = xn.shape
    k = np.arange(xs.shape[1])[None, None, :]
    valid = k < n[None, :, None]
    mask = (xn[:, :, None] <= xs[None, :, :]) & valid
    first = np.where(mask.any(axis=2), mask.argmax(axis=2), 0)
    idx = np.maximum(0, first - 1)
    bi = np.broadcast_to(np.arange(B)[None, :], (E, B))
    x0, x1 = xs[bi, idx], xs[bi, idx + 1]
    y0, y1 = ys[bi, idx], ys[bi, idx + 1]
    return x0, x1, y0, y1


class OracleEnv:
    def __init__(self, spec: S.DistrictSpec, num_envs: int = 1, central_agent=None, stale_observations: bool = True,
                 reward=None, libm_pow: bool = False):
        # The reference takes the round-trip efficiency as `efficiency ** 0.5` on Python / NumPy scalars, i.e. libm `pow`
        # (energy_model.py:675-678), which is not correctly rounded: about once in 10^4 values it differs from sqrt by one ulp,
        # and when the affected float64 then sits on a float32 rounding tie the stored energy balance moves by a float32 ulp.
        # `libm_pow=True` reproduces that with math.pow per element (slow; used when pinning against the reference's traces);
        # the default is the correctly rounded sqrt, which is what the CUDA kernel computes.
        self.libm_pow = bool(libm_pow)
        self.spec = spec
        self.E = int(num_envs)
        self.B = spec.n_buildings
        self.p = spec.params.astype(f64)           # [B, NPARAM]
        self.ip = spec.iparams
        self.table = spec.table                    # [N, W] float32
        self.central = spec.central_agent if central_agent is None else central_agent
        self.stale = stale_observations
        self.entries, self.desc = S.observation_layout(spec, self.central, stale_observations)
        self.flags = self.ip[:, IP['FLAGS']]
        self.reward = reward if reward is not None else resolve_reward(spec)
        self.t = 0
        self.start = np.zeros(self.E, dtype=np.int64)
        self.outage = np.zeros((self.B, 1), dtype=np.float32)
        # building.py:2554: solar_generation = pv.get_generation(series) * -1 stays float64 in the reference
        self.solar64_series = np.stack([S.pv_generation(b, b.series['solar_generation']) * -1 for b in spec.buildings], axis=1)
        self.dyn_weights = []
        for bi, b in enumerate(spec.buildings):
            self.dyn_weights.append(b.dynamics_weights if b.dynamics else None)
        self.evd = spec.ev if getattr(spec, 'ev', None) and (len(spec.ev.get('chargers', [])) or len(spec.ev.get('wms', []))) else None
        if self.evd is not None:
            # charging-constraint state (headroom per limit, violation): set by every apply_actions, NOT touched by reset (building.py:882-892)
            lim = self.evd.get('cc_limits', np.zeros((0, 5)))
            self.cc_state = np.zeros((num_envs, len(lim) * S.CC_SLOTS))
            for k in range(len(lim)):
                self.cc_state[:, k * S.CC_SLOTS:k * S.CC_SLOTS + lim.shape[1]] = lim[k]

    def _root(self, x):
        x = np.asarray(x, dtype=f64)
        if not self.libm_pow:
            return np.sqrt(x)
        import math
        return np.frompyfunc(lambda v: math.pow(v, 0.5) if v >= 0 else float('nan'), 1, 1)(x).astype(f64)

    # -- helpers ---------------------------------------------------------------------------------
    def P(self, name):
        return self.p[:, P[name]][None, :]          # [1, B] broadcast over envs

    def col(self, name, t):
        """value of per-building table column `name` at episode step t for every unit -> [E, B] float64"""
        c = self.ip[:, IP[name]]
        rows = (self.start + t)[:, None]
        return self.table[rows, c[None, :]].astype(f64)

    def col32(self, name, t):
        c = self.ip[:, IP[name]]
        rows = (self.start + t)[:, None]
        return self.table[rows, c[None, :]]

    def solar64(self, t):
        rows = (self.start + t)[:, None]
        return self.solar64_series[rows, np.arange(self.B)[None, :]]

    def flag(self, bit):
        return ((self.flags & bit) != 0)[None, :]

    def cop_cool(self, T):
        # citylearn/energy_model.py:239-250; numerator rounded to float32 first (weak python scalar / float32 array)
        T = w32(T)
        with np.errstate(divide='ignore', invalid='ignore'):
            cop = w32(self.P('CD_COP_NUM')) / (T - w32(self.P('CD_TARGET')))
        return np.where((cop < 0) | (cop > 20), np.float32(20.0), cop).astype(np.float32)

    def eff_heat(self, which, T):
        """COP of a heat-pump heating/dhw device, else the heater efficiency."""
        pre = {'heating': 'HD', 'dhw': 'DD'}[which]
        bit = {'heating': S.F_HEATING_IS_HEAT_PUMP, 'dhw': S.F_DHW_IS_HEAT_PUMP}[which]
        T = w32(T)
        with np.errstate(divide='ignore', invalid='ignore'):
            cop = w32(self.P(f'{pre}_COP_NUM')) / (w32(self.P(f'{pre}_TARGET')) - T)
        cop = np.where((cop < 0) | (cop > 20), np.float32(20.0), cop).astype(f64)
        # heat pump: float32 COP; heater: python-float efficiency (kept in double)
        return np.where(self.flag(bit), cop, self.P(f'{pre}_EFFICIENCY'))

    # -- reset -----------------------------------------------------------------------------------
    def reset(self, episode_start=None, episode_time_steps=None):
        """`CityLearnEnv.reset` (citylearn/citylearn.py:1829-1886) for every env; returns obs [E, L]."""
        spec = self.spec
        E, B = self.E, self.B
        if episode_start is None:
            episode_start = spec.simulation_start_time_step
        self.start = np.broadcast_to(np.asarray(episode_start, dtype=np.int64), (E,)).copy()
        if episode_time_steps is None:
            episode_time_steps = spec.simulation_end_time_step - int(self.start[0]) + 1
        self.T = int(episode_time_steps)
        self.outage = S.outage_signals(spec, self.T, int(self.start[0]))
        self.t = 0
        shape = (E, B)
        self.soc_b = np.broadcast_to(r32(self.P('BAT_INITIAL_SOC')), shape).copy()
        self.cap_deg = np.broadcast_to(self.P('BAT_CAPACITY'), shape).copy()
        self.eff_b = np.broadcast_to(self.P('BAT_EFFICIENCY0'), shape).copy()
        self.cap_deg_is_f32 = False     # python float until the first degrade() (energy_model.py:1055-1056)
        self.eff_is_weak = True         # python float until the first get_current_efficiency()
        self.soc_cs = np.broadcast_to(r32(self.P('CS_INITIAL_SOC')), shape).copy()
        self.soc_hs = np.broadcast_to(r32(self.P('HS_INITIAL_SOC')), shape).copy()
        self.soc_ds = np.broadcast_to(r32(self.P('DS_INITIAL_SOC')), shape).copy()
        # LSTM state (citylearn/dynamics.py:112-127)
        self.has_dyn = bool((self.flags & S.F_DYNAMICS).any())
        if self.has_dyn:
            H = int(self.ip[:, IP['DYN_HIDDEN']].max())
            L = int(self.ip[:, IP['DYN_LOOKBACK']].max())
            nin = int(self.ip[:, IP['DYN_N_INPUTS']].max())
            self.h = np.zeros((E, B, 2, H), dtype=np.float32)
            self.c = np.zeros((E, B, 2, H), dtype=np.float32)
            self.window = np.zeros((E, B, nin, L + 1), dtype=np.float64)
            self.window_fill = 0
            self.cool_dem_override = {}   # cooling demand written by update_cooling_demand at past steps is not needed again
        if self.evd is not None:
            self._ev_reset()
        # t = 0 accumulation from reset -> update_variables (citylearn/building.py:2618-2652)
        dyn = self._time0_values()
        return self._observations(0, dyn, zero_dyn=False)

    def _ec0(self):
        """electricity consumption (float32) added at t == 0 by `update_variables` before any action (building.py:2618-2652)."""
        T = self.col32('C_T_OUT', 0)
        cop_c = self.cop_cool(T)
        ec_cool = self.col32('C_COOLING_DEMAND', 0) / cop_c
        # quirk: heater-type heating device uses dhw_device.get_input_power (building.py:2632)
        hd_eff = np.where(self.flag(S.F_HEATING_IS_HEAT_PUMP), self.eff_heat('heating', T), self.eff_heat('dhw', T))
        # np.array(float32) / python-float efficiency and float32 / float32 COP are both float32 divisions
        ec_heat = self.col32('C_HEATING_DEMAND', 0) / w32(hd_eff)
        ec_dhw = self.col32('C_DHW_DEMAND', 0) / w32(self.eff_heat('dhw', T))
        ec_nsl = self.col32('C_NSL', 0)
        return [np.broadcast_to(x, (self.E, self.B)).astype(np.float32) for x in (ec_cool, ec_heat, ec_dhw, ec_nsl)]

    def _net(self, ec, solar64, outage, chargers32=None, machines32=None):
        """building.py:2685-2703: float64 adds (series getter * np.float64 ratio), + float64 solar (+ the float32 totals of the
        building's chargers and washing machines); stores are float32."""
        r = self.P('TIME_STEP_RATIO')           # np.float64 in the reference -> the getter yields float64 series
        s = ec['cool'].astype(f64) * r + ec['heat'].astype(f64) * r
        s = s + ec['dhw'].astype(f64) * r
        s = s + ec['nsl'].astype(f64) * r
        s = s + ec['bat'].astype(f64) * r
        s = s + solar64
        if chargers32 is not None:
            s = s + chargers32.astype(f64)
            s = s + machines32.astype(f64)
        net64 = np.where(outage, 0.0, s)
        return net64

    def _time0_values(self):
        E, B = self.E, self.B
        ec_cool, ec_heat, ec_dhw, ec_nsl = self._ec0()
        ec = {'cool': ec_cool, 'heat': ec_heat, 'dhw': ec_dhw, 'nsl': ec_nsl, 'bat': np.zeros((E, B), dtype=np.float32)}
        out = (self.outage[:, 0][None, :] > 0) & self.flag(S.F_SIMULATE_OUTAGE)
        net64 = self._net(ec, self.solar64(0), out)
        r = self.P('TIME_STEP_RATIO')
        dyn = np.zeros((E, B, NDYN))
        dyn[..., DYN['electrical_storage_soc']] = self.soc_b
        dyn[..., DYN['cooling_storage_soc']] = self.soc_cs
        dyn[..., DYN['heating_storage_soc']] = self.soc_hs
        dyn[..., DYN['dhw_storage_soc']] = self.soc_ds
        dyn[..., DYN['net_electricity_consumption']] = r32(net64)
        dyn[..., DYN['cooling_electricity_consumption']] = ec_cool * r
        dyn[..., DYN['heating_electricity_consumption']] = ec_heat * r
        dyn[..., DYN['dhw_electricity_consumption']] = ec_dhw * r
        dyn[..., DYN['non_shiftable_load_electricity_consumption']] = ec_nsl * r
        dyn[..., DYN['cooling_demand']] = self.col('C_COOLING_DEMAND', 0)
        dyn[..., DYN['heating_demand']] = self.col('C_HEATING_DEMAND', 0)
        dyn[..., DYN['dhw_demand']] = self.col('C_DHW_DEMAND', 0)
        dyn[..., DYN['indoor_dry_bulb_temperature']] = self.col('C_T_IN', 0)
        dyn[..., DYN['net_electricity_consumption_cost']] = r32(net64 * self.col('C_PRICE', 0))
        dyn[..., DYN['net_electricity_consumption_emission']] = r32(np.maximum(0.0, net64 * self.col('C_CARBON', 0)))
        dyn[..., DYN['electrical_storage_degraded_capacity']] = self.cap_deg
        dyn[..., DYN['energy_to_non_shiftable_load']] = self.col('C_NSL', 0)
        dyn[..., DYN['cooling_demand_series']] = self.col('C_COOLING_DEMAND', 0)
        dyn[..., DYN['heating_demand_series']] = self.col('C_HEATING_DEMAND', 0)
        return dyn

    # -- storage primitives ----------------------------------------------------------------------
    def _energy_init64(self, soc32, cap, loss):
        """`StorageDevice.energy_init` (energy_model.py:661-666): (np.float32 * python float) -> float32 product, then
        times (1 - loss * time_step_ratio) which is np.float64 because `time_step_ratio` is (data.py:428-455)."""
        r = self.P('TIME_STEP_RATIO')
        return np.maximum(0.0, (soc32 * w32(cap)).astype(f64) * (1 - loss * r))

    def _tank_charge(self, pre, soc32, energy64):
        """`StorageTank.charge` -> `StorageDevice.charge` (citylearn/energy_model.py:719-768, 850-870).

        `energy * time_step_ratio` makes everything np.float64; soc / energy_balance are float32 stores.
        """
        r = self.P('TIME_STEP_RATIO')
        cap, eff, loss = self.P(f'{pre}_CAPACITY'), self.P(f'{pre}_EFFICIENCY'), self.P(f'{pre}_LOSS')
        energy64 = energy64 * r                        # StorageTank.charge :863
        bit_in = {'CS': S.F_CS_HAS_MAX_IN, 'HS': S.F_HS_HAS_MAX_IN, 'DS': S.F_DS_HAS_MAX_IN}[pre]
        bit_out = {'CS': S.F_CS_HAS_MAX_OUT, 'HS': S.F_HS_HAS_MAX_OUT, 'DS': S.F_DS_HAS_MAX_OUT}[pre]
        lim_in = self.flag(bit_in) & (energy64 >= 0)
        lim_out = self.flag(bit_out) & (energy64 < 0)
        energy64 = np.where(lim_in, np.fmin(energy64, self.P(f'{pre}_MAX_IN')), energy64)
        energy64 = np.where(lim_out, np.fmax(-self.P(f'{pre}_MAX_OUT'), energy64), energy64)
        energy64 = energy64 * r                        # StorageDevice.charge :732
        e_init = self._energy_init64(soc32, cap, loss)
        rte = self._root(eff)
        fin = np.where(energy64 >= 0, np.minimum(e_init + energy64 * rte, cap), np.maximum(0.0, e_init + energy64 / rte))
        soc = w32(fin / np.maximum(cap, EPS))
        d = fin - e_init
        eb = w32(np.where(d >= 0, d / rte, d * rte))
        return soc, eb

    def _battery_charge(self, energy64, ec_bat32):
        """`Battery.charge` (citylearn/energy_model.py:1027-1141) with the reference's dtype flow.

        State: soc_b (float32 values), cap_deg (float64), eff_b (python float before the first call, np.float64 after).
        """
        r = self.P('TIME_STEP_RATIO')
        cap, pnom = self.P('BAT_CAPACITY'), self.P('BAT_NOMINAL_POWER')
        soc32 = w32(self.soc_b)
        energy64 = energy64 * r
        action_energy = energy64
        e_init = self._energy_init64(soc32, cap, self.P('BAT_LOSS'))
        soc_n = e_init / np.maximum(cap, EPS)                                                 # :1080
        x0, x1, y0, y1 = _interp_curve(soc_n, self.p[:, P['CP_X0']:P['CP_X0'] + S.MAX_CURVE],
                                       self.p[:, P['CP_Y0']:P['CP_Y0'] + S.MAX_CURVE], self.ip[:, IP['CP_N']])
        with np.errstate(divide='ignore', invalid='ignore'):
            p_max = pnom * (y0 + (y1 - y0) * (soc_n - x0) / (x1 - x0))
        # charge branch (:1040-1043)
        avail = pnom - ec_bat32.astype(f64) * r
        e_chg = np.minimum(np.minimum(np.minimum(p_max, avail), self.cap_deg - e_init), energy64)
        arg_chg = np.minimum(action_energy, p_max)
        # discharge branch (:1046-1052): float32 soc difference, PREVIOUS efficiency
        diff32 = soc32 - w32(1.0 - self.P('BAT_DOD'))
        if self.eff_is_weak:
            lim = (diff32 * w32(cap) * w32(self._root(self.eff_b))).astype(f64)
        else:
            lim = (diff32 * w32(cap)).astype(f64) * self._root(self.eff_b)
        lim = -np.maximum(lim, 0.0)
        e_dis = np.maximum(np.maximum(-p_max, lim), energy64)
        arg_dis = np.minimum(np.abs(action_energy), p_max)
        pos = energy64 >= 0
        e = np.where(pos, e_chg, e_dis)
        arg = np.where(pos, arg_chg, arg_dis)
        xn = np.abs(arg) / np.maximum(pnom, EPS)
        x0, x1, y0, y1 = _interp_curve(xn, self.p[:, P['PE_X0']:P['PE_X0'] + S.MAX_CURVE], self.p[:, P['PE_Y0']:P['PE_Y0'] + S.MAX_CURVE], self.ip[:, IP['PE_N']])
        with np.errstate(divide='ignore', invalid='ignore'):
            eff = y0 + (xn - x0) * (y1 - y0) / (x1 - x0)
        # StorageDevice.charge with the new efficiency (:719-768)
        e = e * r
        rte = self._root(eff)
        fin = np.where(e >= 0, np.minimum(e_init + e * rte, cap), np.maximum(0.0, e_init + e / rte))
        soc = w32(fin / np.maximum(cap, EPS))
        d = fin - e_init
        eb32 = w32(np.where(d >= 0, d / rte, d * rte))
        # degrade (:1130-1141): (python * python) * np.float32 -> float32; divided by python float the first time
        # (float32), by np.float64 afterwards; times np.float64 time_step_ratio
        ceb32 = w32(self.P('BAT_CLC') * cap) * np.abs(eb32)
        if self.cap_deg_is_f32:
            deg = ceb32.astype(f64) / (2 * np.maximum(self.cap_deg, EPS)) * r
        else:
            deg = (ceb32 / w32(2 * np.maximum(cap, EPS))).astype(f64) * r
        new_cap = np.maximum(self.cap_deg - deg, 0.0)
        return soc, eb32, eff, new_cap

    # -- one step --------------------------------------------------------------------------------
    def step(self, actions):
        """actions: [E, action_dim] -> (obs [E, L] at t+1, reward [E, R], district [E, 3], dyn [E, B, NDYN] at t)."""
        E, B, t = self.E, self.B, self.t
        a = np.asarray(actions, dtype=np.float32).astype(f64).reshape(E, -1)
        r = self.P('TIME_STEP_RATIO')
        hstep = self.P('HOURS_PER_STEP')

        def act(name, inactive):
            slot = self.ip[:, IP[name]]
            v = a[:, np.maximum(slot, 0)]
            return np.where(slot[None, :] >= 0, v, inactive)

        a_cd = act('A_COOLING_DEVICE', np.nan)
        a_hd = act('A_HEATING_DEVICE', np.nan)
        slot_coh = self.ip[:, IP['A_COOLING_OR_HEATING_DEVICE']]
        if (slot_coh >= 0).any():   # building.py:1550-1553
            coh = act('A_COOLING_OR_HEATING_DEVICE', 0.0)
            a_cd = np.where(slot_coh[None, :] >= 0, np.abs(np.minimum(coh, 0.0)), a_cd)
            a_hd = np.where(slot_coh[None, :] >= 0, np.abs(np.maximum(coh, 0.0)), a_hd)
        a_cs = act('A_COOLING_STORAGE', 0.0)
        a_hs = act('A_HEATING_STORAGE', 0.0)
        a_ds = act('A_DHW_STORAGE', 0.0)
        a_es = act('A_ELECTRICAL_STORAGE', 0.0)

        T_out = self.col32('C_T_OUT', t)
        nsl32 = self.col32('C_NSL', t)
        solar64 = self.solar64(t)
        dem32 = {'cool': self.col32('C_COOLING_DEMAND', t), 'heat': self.col32('C_HEATING_DEMAND', t),
                 'dhw': self.col32('C_DHW_DEMAND', t)}
        dem32 = {k: np.broadcast_to(v, (E, B)).astype(np.float32) for k, v in dem32.items()}
        hvac = self.col('C_HVAC_MODE', t)
        outage = (self.outage[:, t][None, :] > 0) & self.flag(S.F_SIMULATE_OUTAGE)
        outage = np.broadcast_to(outage, (E, B))
        eff = {'cool': self.cop_cool(T_out), 'heat': self.eff_heat('heating', T_out), 'dhw': self.eff_heat('dhw', T_out)}
        # is the efficiency a float32 COP (heat pump) or a python-float heater efficiency?
        eff_is_cop = {'cool': np.ones((1, B), dtype=bool), 'heat': self.flag(S.F_HEATING_IS_HEAT_PUMP), 'dhw': self.flag(S.F_DHW_IS_HEAT_PUMP)}
        pre_dev = {'cool': 'CD', 'heat': 'HD', 'dhw': 'DD'}

        z32 = lambda: np.zeros((E, B), dtype=np.float32)   # noqa: E731
        if t == 0:
            ec_cool, ec_heat, ec_dhw, ec_nsl = self._ec0()
        else:
            ec_cool, ec_heat, ec_dhw, ec_nsl = z32(), z32(), z32(), z32()
        ec = {'cool': ec_cool, 'heat': ec_heat, 'dhw': ec_dhw, 'nsl': ec_nsl, 'bat': z32()}
        eb = {'cs': z32(), 'hs': z32(), 'ds': z32(), 'bat': z32()}
        soc_t = {'cs': w32(self.soc_cs), 'hs': w32(self.soc_hs), 'ds': w32(self.soc_ds)}
        soc_new = dict(soc_t)
        e_from = {k: v.copy() for k, v in dem32.items()}          # energy_from_*_device starts as the demand series (building.py:2555-2557)
        sk_of = {'cool': 'cs', 'heat': 'hs', 'dhw': 'ds'}

        def flex64():
            s = ec['cool'].astype(f64) * r + ec['heat'].astype(f64) * r
            s = s + ec['dhw'].astype(f64) * r
            s = s + ec['nsl'].astype(f64) * r
            s = s + ec['bat'].astype(f64) * r
            cap = np.abs(solar64) - s
            return np.where(outage, np.maximum(0.0, cap), np.inf)

        def add_ec(key, value, mask):
            # `arr[t] += x`: promoted sum, float32 store
            ec[key] = np.where(mask, w32(ec[key].astype(f64) + np.asarray(value, dtype=f64)), ec[key]).astype(np.float32)

        def avail64(key):
            # nominal_power - electricity_consumption[t]; the getter multiplies by the np.float64 time_step_ratio
            return self.P(f'{pre_dev[key]}_NOMINAL_POWER') - ec[key].astype(f64) * r

        def max_output64(key):
            # np.min([python/np.float64, np.float32]) -> np.float64; times float32 COP or python-float efficiency -> float64
            return np.minimum(flex64(), avail64(key)) * eff[key].astype(f64)

        def input_power(key, out64, out_is_f32):
            """`get_input_power`: float32 division when the output is float32, float64 otherwise."""
            e32 = w32(eff[key])
            a32 = (w32(out64) / e32).astype(f64)
            a64 = out64 / eff[key].astype(f64)
            return np.where(out_is_f32, a32, a64)

        def battery(mask):
            if not mask.any():
                return
            energy = np.minimum(a_es * self.P('BAT_NOMINAL_POWER') * hstep, flex64())
            soc, eb32, eff_new, cap_new = self._battery_charge(energy / r, ec['bat'])
            self.soc_b = np.where(mask, soc.astype(f64), self.soc_b)
            self.eff_b = np.where(mask, eff_new, self.eff_b)
            self._pending_cap = np.where(mask, cap_new, self._pending_cap)
            eb['bat'] = np.where(mask, eb32, eb['bat']).astype(np.float32)
            add_ec('bat', eb32, mask)

        self._pending_cap = self.cap_deg.copy()
        dyn_b = self.flag(S.F_DYNAMICS)
        battery_first = a_es < 0.0
        battery(battery_first)
        # dynamics-controlled demand (building.py:3080-3158): first entries of the priority list
        if self.has_dyn and self.window_fill > self.L_max():
            sim = dyn_b & ((self.ip[:, IP['A_COOLING_DEVICE']][None, :] >= 0) | (slot_coh[None, :] >= 0))
            power = a_cd * self.P('CD_NOMINAL_POWER') * hstep
            cd32 = ((self.flags & S.F_CD_NOMINAL_F32) != 0)[None, :]      # autosized: np.float32 nominal power (building.py:3110)
            power = np.where(cd32, ((a_cd.astype(np.float32) * self.P('CD_NOMINAL_POWER').astype(np.float32)) * np.float32(hstep)).astype(f64), power)
            d = np.minimum(power, avail64('cool')) * eff['cool'].astype(f64)
            d = np.where((hvac == 1) | (hvac == 3), d, 0.0)
            dem32['cool'] = np.where(sim, w32(d), dem32['cool']).astype(np.float32)
            simh = dyn_b & ((self.ip[:, IP['A_HEATING_DEVICE']][None, :] >= 0) | (slot_coh[None, :] >= 0))
            powerh = a_hd * self.P('HD_NOMINAL_POWER')          # no hours factor (:3146)
            hd32 = ((self.flags & S.F_HD_NOMINAL_F32) != 0)[None, :]
            powerh = np.where(hd32, (a_hd.astype(np.float32) * self.P('HD_NOMINAL_POWER').astype(np.float32)).astype(f64), powerh)
            dh = np.minimum(powerh, avail64('heat')) * eff['heat'].astype(f64)
            dh = np.where((hvac == 2) | (hvac == 3), dh, 0.0)
            dem32['heat'] = np.where(simh, w32(dh), dem32['heat']).astype(np.float32)
            e_from['cool'] = np.where(sim, e_from['cool'], e_from['cool'])

        def device(key):
            sk = sk_of[key]
            storage_out32 = -np.minimum(eb[sk], np.float32(0.0))
            cand32 = dem32[key] - storage_out32
            mo64 = max_output64(key)
            is32 = cand32.astype(f64) <= mo64                 # python min() keeps the first minimal argument (the float32 one)
            out64 = np.where(is32, cand32.astype(f64), mo64)
            e_from[key] = w32(out64)
            cons = input_power(key, out64, is32 & np.broadcast_to(eff_is_cop[key] | True, (E, B)))
            add_ec(key, np.maximum(0.0, cons), np.ones((E, B), dtype=bool))

        def storage(key, pre_tank, cap_for_action, hours, a_s, mask, cap_f32=None):
            if not mask.any():
                return
            sk = sk_of[key]
            energy = a_s * cap_for_action * hours                # python float ...
            if cap_f32 is not None:                              # ... unless the tank was autosized: np.float32 capacity
                e32 = (a_s.astype(np.float32) * cap_for_action.astype(np.float32)) * np.float32(hours)
                energy = np.where(cap_f32[None, :], e32.astype(f64), energy)
            mo64 = max_output64(key)
            up = energy > 0.0
            lim_up = mo64 <= energy                             # min(max_output, energy): np.float64 wins when smaller or equal
            nd32 = -dem32[key]
            lim_dn = nd32.astype(f64) >= energy                 # max(-demand, energy): np.float32 wins when larger or equal
            e64 = np.where(up, np.where(lim_up, mo64, energy), np.where(lim_dn, nd32.astype(f64), energy))
            soc, ebal = self._tank_charge(pre_tank, soc_t[sk], e64 / r)
            soc_new[sk] = np.where(mask, soc, soc_new[sk]).astype(np.float32)
            eb[sk] = np.where(mask, ebal, eb[sk]).astype(np.float32)
            charged32 = np.maximum(ebal, np.float32(0.0))
            cons = (charged32 / w32(eff[key])).astype(f64)      # float32 / float32 COP, or np.array(float32) / python efficiency
            add_ec(key, cons, mask)

        thermal = bool((self.flags & S.F_HAS_THERMAL).any())
        if thermal:
            # storage goes before its device when discharging (building.py:1614-1622)
            cs32, hs32 = (self.flags & S.F_CS_CAPACITY_F32) != 0, (self.flags & S.F_HS_CAPACITY_F32) != 0
            plan = (('cool', 'CS', self.P('CS_CAPACITY'), 1.0, a_cs, cs32),
                    ('heat', 'HS', self.P('CS_CAPACITY'), hstep, a_hs, cs32),    # COOLING tank capacity (building.py:1720)
                    ('dhw', 'DS', self.P('HS_CAPACITY'), hstep, a_ds, hs32))     # HEATING tank capacity (building.py:1765)
            for key, pre_tank, capa, hours, a_s, c32 in plan:
                neg = np.broadcast_to(a_s < 0.0, (E, B))
                storage(key, pre_tank, capa, hours, a_s, neg, c32)
                device(key)
                storage(key, pre_tank, capa, hours, a_s, ~neg, c32)
            self.soc_cs, self.soc_hs, self.soc_ds = [soc_new[k].astype(f64) for k in ('cs', 'hs', 'ds')]
        # non-shiftable load (building.py:1784-1789)
        dem_nsl = np.minimum(np.broadcast_to(nsl32, (E, B)).astype(f64), flex64())
        e_to_nsl32 = w32(dem_nsl)
        add_ec('nsl', dem_nsl, np.ones((E, B), dtype=bool))
        battery(~np.broadcast_to(battery_first, (E, B)))
        ch_ec32 = wm_ec32 = None
        if self.evd is not None:
            ch_ec32, wm_ec32 = self._ev_step(t, a)      # appended to the priority list (building.py:1582-1604): chargers, then washing machines
        self.cap_deg = self._pending_cap
        self.cap_deg_is_f32 = True
        self.eff_is_weak = False

        # observations at t needed by dynamics and rewards (building.py:1435-1437)
        obs_dem = {k: (e_from[k] + np.abs(np.minimum(eb[sk_of[k]], np.float32(0.0)))).astype(f64) for k in ('cool', 'heat', 'dhw')}
        t_in = np.broadcast_to(self.col('C_T_IN', t), (E, B)).copy()
        if self.has_dyn:
            t_in = self._dynamics(t, obs_dem['cool'], t_in)

        # update_variables (building.py:2615-2703)
        allm = np.ones((E, B), dtype=bool)
        if t == 0:
            add_ec('cool', ((e_from['cool'] + eb['cs']) / w32(eff['cool'])).astype(f64), allm)
            hd_eff0 = np.where(self.flag(S.F_HEATING_IS_HEAT_PUMP), eff['heat'], eff['dhw'])
            add_ec('heat', ((e_from['heat'] + eb['hs']) / w32(hd_eff0)).astype(f64), allm)
            add_ec('dhw', ((e_from['dhw'] + eb['ds']) / w32(eff['dhw'])).astype(f64), allm)
            add_ec('nsl', e_to_nsl32, allm)
            add_ec('bat', eb['bat'], allm)
        net64 = self._net(ec, solar64, outage, ch_ec32, wm_ec32)
        net = r32(net64)
        cost = r32(net64 * self.col('C_PRICE', t))
        emission = r32(np.maximum(0.0, net64 * self.col('C_CARBON', t)))
        # district sums: python sum() of np.float32 in building order (citylearn.py:1908-1918)
        district = np.zeros((E, 3), dtype=np.float32)
        for bi in range(B):
            district[:, 0] = district[:, 0] + w32(net[:, bi])
            district[:, 1] = district[:, 1] + w32(cost[:, bi])
            district[:, 2] = district[:, 2] + w32(emission[:, bi])
        district = district.astype(f64)

        r32w = r
        dyn = np.zeros((E, B, NDYN))
        dyn[..., DYN['electrical_storage_soc']] = self.soc_b
        dyn[..., DYN['cooling_storage_soc']] = self.soc_cs
        dyn[..., DYN['heating_storage_soc']] = self.soc_hs
        dyn[..., DYN['dhw_storage_soc']] = self.soc_ds
        dyn[..., DYN['net_electricity_consumption']] = net
        dyn[..., DYN['cooling_demand']] = obs_dem['cool']
        dyn[..., DYN['heating_demand']] = obs_dem['heat']
        dyn[..., DYN['dhw_demand']] = obs_dem['dhw']
        dyn[..., DYN['cooling_electricity_consumption']] = ec['cool'] * r32w
        dyn[..., DYN['heating_electricity_consumption']] = ec['heat'] * r32w
        dyn[..., DYN['dhw_electricity_consumption']] = ec['dhw'] * r32w
        dyn[..., DYN['cooling_storage_electricity_consumption']] = eb['cs'] / w32(eff['cool'])
        dyn[..., DYN['heating_storage_electricity_consumption']] = eb['hs'] / w32(eff['heat'])
        dyn[..., DYN['dhw_storage_electricity_consumption']] = eb['ds'] / w32(eff['dhw'])
        dyn[..., DYN['electrical_storage_electricity_consumption']] = ec['bat'] * r32w
        dyn[..., DYN['indoor_dry_bulb_temperature']] = t_in
        dyn[..., DYN['non_shiftable_load_electricity_consumption']] = ec['nsl'] * r32w
        dyn[..., DYN['electrical_storage_energy_balance']] = eb['bat']
        dyn[..., DYN['cooling_storage_energy_balance']] = eb['cs']
        dyn[..., DYN['heating_storage_energy_balance']] = eb['hs']
        dyn[..., DYN['dhw_storage_energy_balance']] = eb['ds']
        dyn[..., DYN['net_electricity_consumption_cost']] = cost
        dyn[..., DYN['net_electricity_consumption_emission']] = emission
        dyn[..., DYN['electrical_storage_degraded_capacity']] = self.cap_deg
        dyn[..., DYN['energy_to_non_shiftable_load']] = e_to_nsl32
        dyn[..., DYN['cooling_demand_series']] = dem32['cool']
        dyn[..., DYN['heating_demand_series']] = dem32['heat']
        self.last_aux = {'cooling_demand_series': dem32['cool'], 'heating_demand_series': dem32['heat'],
                         'energy_from_cooling_device': e_from['cool'], 'energy_from_dhw_device': e_from['dhw'],
                         'efficiency': self.eff_b.copy(), 'outage': outage}

        reward = self.reward.calculate(self, t, dyn, district)
        self.t = t + 1
        if self.evd is not None:
            self._ev_advance(self.t)
        if self.stale:
            obs = self._observations(self.t, None, zero_dyn=True)
        else:
            obs = self._observations(self.t, dyn, zero_dyn=False)
        return obs.astype(np.float32), reward.astype(np.float32), district.astype(np.float32), dyn

    # -- electric vehicles, chargers, washing machines ---------------------------
    def _ev_reset(self):
        """`ElectricVehicle.reset` / `Battery.reset` + `associate_chargers_to_electric_vehicles` at t = 0 (citylearn.py:1871-1874)."""
        ev, E = self.evd, self.E
        q = ev['ev_params']
        n = ev['n_ev']
        self.ev_soc_prev = np.zeros((E, n))                                                  # soc[t - 1] entries (float32 values)
        self.ev_soc = np.broadcast_to(r32(q[:, P['BAT_INITIAL_SOC']]), (E, n)).copy()        # soc[t] entries
        self.ev_cap_deg = np.broadcast_to(q[:, P['BAT_CAPACITY']], (E, n)).copy()
        self.ev_eff = np.broadcast_to(q[:, P['BAT_EFFICIENCY0']], (E, n)).copy()
        self.ev_charged = np.zeros((E, n), dtype=bool)          # efficiency / degraded capacity are python floats until the first charge()
        t0 = self.table[self.start[:, None], ev['ev_cols'][:, 3][None, :]].astype(f64)      # every connection is new at t = 0
        self.ev_soc = np.where(np.isnan(t0), self.ev_soc, t0)
        self.wm_initiated = np.zeros((E, len(ev['wms'])), dtype=bool)
        self.last_ev = None

    def _ev_interp(self, x, xs, ys, n):
        """np.interp(x, xs[:n], ys[:n]) element-wise (Charger.get_efficiency, electric_vehicle_charger.py:252-281)."""
        return np.array([np.interp(v, xs[:n], ys[:n]) for v in np.atleast_1d(x)])

    def _ev_battery_charge(self, v, energy64):
        """`Battery.charge(energy)` of vehicle v[e] for every env (energy_model.py:1027-1057): returns (soc32, eb32, eff, cap_deg)."""
        ev = self.evd
        q = ev['ev_params'][v]                                   # [E, NPARAM]
        E = len(v)
        ar = np.arange(E)
        r = q[:, P['TIME_STEP_RATIO']]
        cap, pnom = q[:, P['BAT_CAPACITY']], q[:, P['BAT_NOMINAL_POWER']]
        t = self.t
        soc_init32 = w32(self.ev_soc_prev[ar, v] if t > 0 else self.ev_soc[ar, v])           # soc[t-1], or soc[0] at t == 0 (:664-666, :1046)
        charged = self.ev_charged[ar, v]
        cap_deg, eff_prev = self.ev_cap_deg[ar, v], self.ev_eff[ar, v]
        energy64 = energy64 * r
        action_energy = energy64
        e_init = np.maximum(0.0, (soc_init32 * w32(cap)).astype(f64) * (1 - q[:, P['BAT_LOSS']] * r))
        soc_n = e_init / np.maximum(cap, EPS)

        def seg(xn, x_off, y_off, n):
            xs = q[:, x_off:x_off + S.MAX_CURVE]; ys = q[:, y_off:y_off + S.MAX_CURVE]
            k = np.arange(S.MAX_CURVE)[None, :]
            mask = (xn[:, None] <= xs) & (k < n[:, None])
            first = np.where(mask.any(axis=1), mask.argmax(axis=1), 0)
            idx = np.maximum(0, first - 1)
            return xs[ar, idx], xs[ar, idx + 1], ys[ar, idx], ys[ar, idx + 1]
        n_pe, n_cp = ev['ev_ip'][v, 0], ev['ev_ip'][v, 1]
        x0, x1, y0, y1 = seg(soc_n, P['CP_X0'], P['CP_Y0'], n_cp)
        with np.errstate(divide='ignore', invalid='ignore'):
            p_max = pnom * (y0 + (y1 - y0) * (soc_n - x0) / (x1 - x0))
        avail = pnom - 0.0 * r                                   # the vehicle battery's own consumption at t is 0 before its one charge
        e_chg = np.minimum(np.minimum(np.minimum(p_max, avail), cap_deg - e_init), energy64)
        arg_chg = np.minimum(action_energy, p_max)
        diff32 = soc_init32 - w32(1.0 - q[:, P['BAT_DOD']])
        root_prev = self._root(eff_prev)
        lim = np.where(charged, (diff32 * w32(cap)).astype(f64) * root_prev, (diff32 * w32(cap) * w32(root_prev)).astype(f64))
        lim = -np.maximum(lim, 0.0)
        e_dis = np.maximum(np.maximum(-p_max, lim), energy64)
        arg_dis = np.minimum(np.abs(action_energy), p_max)
        pos = energy64 >= 0
        e = np.where(pos, e_chg, e_dis)
        arg = np.where(pos, arg_chg, arg_dis)
        xn = np.abs(arg) / np.maximum(pnom, EPS)
        x0, x1, y0, y1 = seg(xn, P['PE_X0'], P['PE_Y0'], n_pe)
        with np.errstate(divide='ignore', invalid='ignore'):
            eff = y0 + (xn - x0) * (y1 - y0) / (x1 - x0)
        e = e * r
        rte = self._root(eff)
        fin = np.where(e >= 0, np.minimum(e_init + e * rte, cap), np.maximum(0.0, e_init + e / rte))
        soc = w32(fin / np.maximum(cap, EPS))
        d = fin - e_init
        eb32 = w32(np.where(d >= 0, d / rte, d * rte))
        ceb32 = w32(q[:, P['BAT_CLC']] * cap) * np.abs(eb32)
        deg = np.where(charged, ceb32.astype(f64) / (2 * np.maximum(cap_deg, EPS)) * r, (ceb32 / w32(2 * np.maximum(cap, EPS))).astype(f64) * r)
        return soc, eb32, eff, np.maximum(cap_deg - deg, 0.0)

    def _apply_charging_constraints(self, a, hours):
        """`Building._apply_charging_constraints_to_actions` (citylearn/building.py:894-982) for every constrained building and env:
        python-float arithmetic, sums in dictionary order (`0 + x0 + x1 ...`).  Updates the headroom / violation state the next
        observation reports and returns the (float64) action matrix with the scaled charger actions."""
        ev = self.evd
        ccb = ev.get('cc_building', ())
        if len(ccb) == 0:
            return a
        from citylearn_b200.ev import CHARGER_PARAMS as CP
        a = np.array(a, dtype=f64)
        nph = ev['cc_limits'].shape[1] - 1
        for k, bi in enumerate(ccb):
            ks = [i for i, c in enumerate(ev['chargers']) if c.building == bi]       # the building's chargers, in order
            blim = ev['cc_limits'][k, 0]
            plim = ev['cc_limits'][k, 1:]
            members = [[int(m) for m in ev['cc_members'][k, j] if m >= 0] for j in range(nph)]
            maxp = [float(ev['ch_params'][i][CP['MAX_C']]) for i in ks]
            slots = [int(ev['ch_action'][i]) for i in ks]
            for e in range(self.E):
                req, scale = {}, {}
                for j, sl in enumerate(slots):
                    if sl < 0:
                        continue
                    act = float(a[e, sl])
                    if act <= 0.0 or maxp[j] <= 0.0:
                        continue
                    req[j] = act * maxp[j]
                    scale[j] = 1.0
                state = self.cc_state[e, k * S.CC_SLOTS:(k + 1) * S.CC_SLOTS]
                state[0] = blim
                state[1:1 + nph] = plim
                state[S.CC_SLOTS - 1] = 0.0
                if not req:
                    continue
                viol = 0.0
                total = sum(req.values())
                if not np.isnan(blim) and blim >= 0.0 and total > blim:
                    sc = 0.0 if blim == 0 else blim / total
                    for j in scale:
                        scale[j] *= sc
                    viol += total - blim
                for j_ph in range(nph):
                    lim = plim[j_ph]
                    if np.isnan(lim) or lim < 0.0:
                        continue
                    psum = sum(req.get(m, 0.0) * scale.get(m, 1.0) for m in members[j_ph] if m in req)
                    if psum > lim:
                        ps = 0.0 if lim == 0 else lim / psum
                        for m in members[j_ph]:
                            if m in scale:
                                scale[m] *= ps
                        viol += psum - lim
                scaled = {j: req[j] * scale.get(j, 1.0) for j in req}
                for j, sl in enumerate(slots):
                    if sl < 0:
                        continue
                    act = float(a[e, sl])
                    if act <= 0.0:
                        continue
                    a[e, sl] = 0.0 if maxp[j] <= 0.0 else max(0.0, min(act, scaled.get(j, 0.0) / maxp[j]))
                used = sum(scaled.values())
                state[0] = blim - used
                for j_ph in range(nph):
                    if not np.isnan(plim[j_ph]):
                        state[1 + j_ph] = plim[j_ph] - sum(scaled.get(m, 0.0) for m in members[j_ph])
                state[S.CC_SLOTS - 1] = viol * hours
        return a

    def _ev_step(self, t, a):
        """`Charger.update_connected_electric_vehicle_soc` for every charger, `WashingMachine.start_cycle` for every machine
        (electric_vehicle_charger.py:283-329, energy_model.py:1311-1327); returns the per-building float32 consumption totals."""
        ev, E, B = self.evd, self.E, self.B
        rows = self.start + t
        ar = np.arange(E)
        CH = ev['chargers']
        from citylearn_b200.ev import CHARGER_PARAMS as CP
        nc = len(CH)
        ch_ec = np.zeros((E, nc), dtype=np.float32)
        past = np.zeros((E, nc), dtype=np.float32)
        info = {k: np.zeros((E, nc)) for k in ('connected', 'soc_prev', 'soc_now', 'capacity', 'min_capacity', 'required', 'hours')}
        hours = self.spec.seconds_per_time_step / 3600
        a = self._apply_charging_constraints(a, hours)
        for k, c in enumerate(CH):
            slot = ev['ch_action'][k]
            q = ev['ch_params'][k]
            conn = self.table[rows, ev['ch_cols'][k, 0]] > 0
            v = np.maximum(self.table[rows, ev['ch_cols'][k, 1]].astype(np.int64), 0)
            act = a[:, slot] if slot >= 0 else np.zeros(E)
            nz = (act != 0) & (slot >= 0)
            charging = act > 0
            n_c, n_d = int(q[CP['C_N']]), int(q[CP['D_N']])
            eff_c = self._ev_interp(np.abs(act), q[CP['C_X0']:CP['C_X0'] + 8], q[CP['C_Y0']:CP['C_Y0'] + 8], n_c) if n_c else np.full(E, q[CP['EFF']])
            eff_d = self._ev_interp(np.abs(act), q[CP['D_X0']:CP['D_X0'] + 8], q[CP['D_Y0']:CP['D_Y0'] + 8], n_d) if n_d else np.full(E, q[CP['EFF']])
            eff = np.where(charging, eff_c, eff_d)
            eff_is_f64 = np.where(charging, bool(n_c), bool(n_d))        # np.interp yields np.float64, the flat efficiency is a python float
            en_c = np.maximum(np.minimum(act * q[CP['MAX_C']] * hours, q[CP['MAX_C']]), q[CP['MIN_C']])
            en_d = np.maximum(np.minimum(act * q[CP['MAX_D']] * hours, -q[CP['MIN_D']]), -q[CP['MAX_D']])
            energy = np.where(charging, en_c, en_d)
            with np.errstate(divide='ignore', invalid='ignore'):
                energy_kwh = np.where(charging, energy * eff, energy / eff)
            past[:, k] = np.where(nz, energy, 0.0).astype(np.float32)
            do = nz & conn
            if do.any():
                soc, eb32, eff_new, cap_new = self._ev_battery_charge(v, energy_kwh)
                self.ev_soc[ar[do], v[do]] = soc[do].astype(f64)
                self.ev_eff[ar[do], v[do]] = eff_new[do]
                self.ev_cap_deg[ar[do], v[do]] = cap_new[do]
                self.ev_charged[ar[do], v[do]] = True
                # battery_energy_balance / efficiency (>= 0) or * efficiency: float32 arithmetic with the python-float efficiency,
                # float64 with an interpolated one; stored into a float32 array either way
                e32 = w32(eff)
                with np.errstate(divide='ignore', invalid='ignore'):
                    c32 = np.where(eb32 >= 0, eb32 / e32, eb32 * e32)
                    c64 = w32(np.where(eb32 >= 0, eb32.astype(f64) / eff, eb32.astype(f64) * eff))
                ch_ec[:, k] = np.where(do, np.where(eff_is_f64, c64, c32), 0.0).astype(np.float32)
            info['connected'][:, k] = conn
            info['soc_prev'][:, k] = np.where(t == 0, ev['ev_params'][v, P['BAT_INITIAL_SOC']], self.ev_soc_prev[ar, v])
            info['soc_now'][:, k] = self.ev_soc[ar, v]
            info['capacity'][:, k] = ev['ev_params'][v, P['BAT_CAPACITY']]
            info['min_capacity'][:, k] = (1 - ev['ev_params'][v, P['BAT_DOD']]) * ev['ev_params'][v, P['BAT_CAPACITY']]
            info['required'][:, k] = self.table[rows, ev['ch_cols'][k, 2]].astype(f64)
            info['hours'][:, k] = self.table[rows, ev['ch_cols'][k, 3]].astype(f64)
        info['past'] = past
        info['ch_ec'] = ch_ec
        # washing machines
        WM = ev['wms']
        wm_ec = np.zeros((E, len(WM)), dtype=np.float32)
        for k, w in enumerate(WM):
            slot = ev['wm_action'][k]
            st = self.table[rows, ev['wm_cols'][k, 0]].astype(np.int64)
            en = self.table[rows, ev['wm_cols'][k, 1]].astype(np.int64)
            if t > 0:      # WashingMachine.next_time_step (energy_model.py:1302-1309): a new window re-arms the machine
                pst = self.table[rows - 1, ev['wm_cols'][k, 0]].astype(np.int64)
                pen = self.table[rows - 1, ev['wm_cols'][k, 1]].astype(np.int64)
                self.wm_initiated[:, k] &= ~((pst != st) | (pen != en))
            if slot < 0:
                continue
            act = a[:, slot]
            go = (~self.wm_initiated[:, k]) & (act > 0) & (st != -1) & (en != -1) & (st <= t) & (t <= en)
            load = w.profile_sum[rows]
            # every entry of the profile whose step t + offset lies inside the episode is ADDED to ec[t] (float32 accumulate)
            for e_i in np.nonzero(go)[0]:
                acc = np.float32(0.0)
                for off in range(int(w.profile_len[rows[e_i]])):
                    if t + off < self.T:
                        acc = np.float32(np.float64(acc) + np.float64(w.profile_values(rows[e_i])[off]))    # float32 slot += np.float64 entry
                wm_ec[e_i, k] = acc
            self.wm_initiated[:, k] |= go & (w.profile_len[rows] > 0)
        info['wm_ec'] = wm_ec
        self.last_ev = info
        # building totals: python `0 + float32 + ...` in charger / machine order (building.py:2654-2672) -> float32
        ch_tot = np.zeros((E, B), dtype=np.float32)
        for k, c in enumerate(CH):
            ch_tot[:, c.building] = (ch_tot[:, c.building] + ch_ec[:, k]).astype(np.float32)
        wm_tot = np.zeros((E, B), dtype=np.float32)
        for k, w in enumerate(WM):
            wm_tot[:, w.building] = (wm_tot[:, w.building] + wm_ec[:, k]).astype(np.float32)
        return ch_tot, wm_tot

    def _ev_advance(self, t_new):
        """`next_time_step` for the vehicles: new soc[t] entry = 0, then `simulate_unconnected_ev_soc` and
        `associate_chargers_to_electric_vehicles` (citylearn.py:1336-1351) - compiled per row by ev.compile_schedule."""
        ev = self.evd
        self.ev_soc_prev = self.ev_soc.copy()
        self.ev_soc = np.zeros_like(self.ev_soc)
        if t_new > self.T - 1:
            return
        rows = self.start + t_new
        cols = ev['ev_cols']
        assoc = self.table[rows[:, None], cols[:, 0][None, :]].astype(f64)
        if t_new + 1 < self.T:                                   # simulate_unconnected_ev_soc returns early on the last step (:1415-1417)
            sim = self.table[rows[:, None], cols[:, 1][None, :]].astype(f64)
            drift = ev['schedule']['drift'][rows]                # float64 factors (the table is float32)
            drifted = r32(np.clip(self.ev_soc_prev * drift, 0.0, 1.0))
            self.ev_soc = np.where(np.isnan(drift), self.ev_soc, drifted)
            self.ev_soc = np.where(np.isnan(sim), self.ev_soc, sim)
        self.ev_soc = np.where(np.isnan(assoc), self.ev_soc, assoc)

    # -- LSTM dynamics ---------------------------------------------------------------------------
    def L_max(self):
        return int(self.ip[:, IP['DYN_LOOKBACK']].max())

    def _dynamics(self, t, obs_cool_dem, t_in):
        """`_update_dynamics_input` + `update_indoor_dry_bulb_temperature` (citylearn/building.py:3000-3078)."""
        spec = self.spec
        E, B = self.E, self.B
        L = self.L_max()
        t_in = t_in.copy()
        new = np.zeros((E, B, self.window.shape[2]))
        for bi, b in enumerate(spec.buildings):
            if not b.dynamics:
                continue
            a = b.dynamics_attrs
            rows = self.start + t
            for i, (k, mn, mx) in enumerate(zip(a['input_observation_names'], a['input_normalization_minimum'], a['input_normalization_maximum'])):
                if k == 'indoor_dry_bulb_temperature':
                    v = t_in[:, bi]
                elif k == 'cooling_demand':
                    v = obs_cool_dem[:, bi]
                elif k.endswith('_sin') or k.endswith('_cos'):
                    base = k[:-4]
                    x = 2 * np.pi * b.series[base][rows] / S.PERIODIC[base]
                    v = np.sin(x) if k.endswith('_sin') else np.cos(x)
                else:
                    v = b.series[k][rows].astype(f64)
                    v = r32(v)
                new[:, bi, i] = (v - mn) / (mx - mn)
        self.window = np.concatenate([self.window[..., 1:], new[..., None]], axis=-1)
        self.window_fill += 1
        if self.window_fill <= L:      # W[0][0] is still None (building.py:2996-2998)
            return t_in
        for bi, b in enumerate(spec.buildings):
            if not b.dynamics:
                continue
            a = b.dynamics_attrs
            ix = a['input_observation_names'].index('indoor_dry_bulb_temperature')
            w = b.dynamics_weights
            X = np.empty((E, L, self.window.shape[2]), dtype=np.float32)
            for i in range(self.window.shape[2]):
                X[:, :, i] = self.window[:, bi, i, :-1] if i == ix else self.window[:, bi, i, 1:]
            h = self.h[:, bi].copy()
            c = self.c[:, bi].copy()
            y = None
            for s in range(L):
                inp = X[:, s, :]
                for layer in range(2):
                    gates = (inp @ w[f'l_lstm.weight_ih_l{layer}'].T + w[f'l_lstm.bias_ih_l{layer}']
                             + h[:, layer] @ w[f'l_lstm.weight_hh_l{layer}'].T + w[f'l_lstm.bias_hh_l{layer}']).astype(np.float32)
                    H = h.shape[-1]
                    i_g = _sigmoid(gates[:, 0:H]); f_g = _sigmoid(gates[:, H:2 * H])
                    g_g = np.tanh(gates[:, 2 * H:3 * H]); o_g = _sigmoid(gates[:, 3 * H:4 * H])
                    c[:, layer] = (f_g * c[:, layer] + i_g * g_g).astype(np.float32)
                    h[:, layer] = (o_g * np.tanh(c[:, layer])).astype(np.float32)
                    inp = h[:, layer]
            y = (h[:, 1] @ w['l_linear.weight'].T + w['l_linear.bias']).astype(np.float32)[:, 0]
            self.h[:, bi], self.c[:, bi] = h, c
            self.window[:, bi, ix, -1] = y.astype(f64)                 # normalised prediction overwrites the slot (:3027-3028)
            lo, hi = a['input_normalization_minimum'][ix], a['input_normalization_maximum'][ix]
            t_in[:, bi] = r32((y * np.float32(hi - lo) + np.float32(lo)).astype(np.float32))   # torch float32 arithmetic (:3031-3037)
        return t_in

    # -- observations ----------------------------------------------------------------------------
    def _observations(self, t, dyn, zero_dyn):
        E = self.E
        L = len(self.desc)
        obs = np.zeros((E, L), dtype=np.float64)
        t_eff = min(t, self.T - 1)
        rows = self.start + t_eff
        for j, (kind, a, b_, bi) in enumerate(self.desc):
            if kind == S.OBS_TS:
                obs[:, j] = self.table[rows, b_ - 1 if (t == 0 and b_ > 0) else a]      # b: the column to read at t = 0 (+ 1), if any
            elif kind == S.OBS_DYN:
                obs[:, j] = 0.0 if zero_dyn else dyn[:, bi, a]
            elif kind == S.OBS_OUTAGE:
                obs[:, j] = self.outage[bi, t_eff]
            elif kind == S.OBS_STATE:              # charging-constraint headroom / violation of the last applied actions (not reset)
                obs[:, j] = self.cc_state[:, a]
        return obs


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# reward functions (citylearn/reward_function.py)
# ------------------------------------------------------------------------------------------------
class OracleReward:
    """Built-in reward evaluated from the post-action state at t (citylearn/citylearn.py:1022-1023)."""

    def __init__(self, kind='RewardFunction', **attrs):
        self.kind = kind
        self.attrs = attrs

    def _comfort(self, env, t, dyn):
        """ComfortReward (reward_function.py:269-334); observation values are np.float32, attributes Python floats -> float32."""
        band_attr = self.attrs.get('band')
        lo_e = np.float32(2.0 if self.attrs.get('lower_exponent') is None else self.attrs['lower_exponent'])
        hi_e = np.float32(2.0 if self.attrs.get('higher_exponent') is None else self.attrs['higher_exponent'])
        heating = dyn[..., DYN['heating_demand']] > dyn[..., DYN['cooling_demand']]
        mode = env.col('C_HVAC_MODE', t)
        T = w32(dyn[..., DYN['indoor_dry_bulb_temperature']])
        csp, hsp = env.col32('C_COOL_SP', t), env.col32('C_HEAT_SP', t)
        band = env.col32('C_COMFORT_BAND', t) if band_attr is None else np.float32(band_attr)
        with np.errstate(invalid='ignore'):
            # hvac_mode in [1, 2]
            sp = np.where(mode == 1, csp, hsp).astype(np.float32)
            lower, upper = sp - band, sp + band
            delta = np.abs(T - sp)
            r_a = -delta ** np.where(mode == 2, lo_e, hi_e)
            r_b = np.where(heating, np.float32(0.0), -delta)
            r_c = np.where(heating, -delta, np.float32(0.0))
            r_d = -delta ** np.where(heating, hi_e, lo_e)
            r12 = np.where(T < lower, r_a, np.where((lower <= T) & (T < sp), r_b, np.where((sp <= T) & (T <= upper), r_c, r_d)))
            # other modes
            lower, upper = hsp - band, csp + band
            cd, hd = T - csp, T - hsp
            r_a = -np.abs(hd) ** np.where(~heating, hi_e, lo_e)
            r_b = -np.abs(hd)
            r_d = -np.abs(cd)
            r_e = -np.abs(cd) ** np.where(heating, hi_e, lo_e)
            r03 = np.where(T < lower, r_a, np.where((lower <= T) & (T < hsp), r_b, np.where((hsp <= T) & (T <= csp), np.float32(0.0),
                           np.where((csp < T) & (T < upper), r_d, r_e))))
        return np.where((mode == 1) | (mode == 2), r12, r03).astype(np.float32)

    def _solar_penalty(self, env, dyn):
        """SolarPenaltyReward (reward_function.py:189-214): float32 observation arithmetic."""
        e = w32(dyn[..., DYN['net_electricity_consumption']])
        reward = np.zeros_like(e)
        for pre, key in (('CS', 'cooling_storage_soc'), ('HS', 'heating_storage_soc'), ('DS', 'dhw_storage_soc'), ('BAT', 'electrical_storage_soc')):
            cap = env.P(f'{pre}_CAPACITY')
            s = w32(dyn[..., DYN[key]])
            term = -(np.float32(1.0) + np.sign(e) * s) * np.abs(e)
            reward = reward + np.where(cap > EPS, term, np.float32(0.0)).astype(np.float32)
        return reward

    def _ev_reward(self, env, t, dyn, district):
        """Electric_Vehicles_Reward_Function (reward_function.py:389-523): the MARL reward only scales the penalty / bonus terms of the
        building's chargers; a building without chargers is rewarded 0."""
        E, B = dyn.shape[0], dyn.shape[1]
        w = self.attrs.get('weights') or {"no_car_charging": -5.0, "battery_limits": -2.0, "soc_impossible": -10.0, "soc_under": -5.0,
                                          "close_soc": 10.0, "self_ev_consumption": 5.0, "extra_self_production": 5.0}
        e = dyn[..., DYN['net_electricity_consumption']]
        be = e * -1
        marl = np.sign(be) * 0.01 * be ** 2 * np.fmax(0.0, district[:, 0:1])
        if env.central:
            tot = np.zeros(E)
            for bi in range(B):
                tot = tot + marl[:, bi]
            marl = np.broadcast_to(tot[:, None], (E, B))
        out = np.zeros((E, B))
        info = env.last_ev
        from citylearn_b200.ev import CHARGER_PARAMS as CP
        for k, c in enumerate(env.evd['chargers']):
            bi = c.building
            mult = 1.0 / (1.0 + np.abs(marl[:, bi]))
            q = env.evd['ch_params'][k]
            con = info['connected'][:, k] > 0
            kwh = info['past'][:, k].astype(f64)
            net = w32(e[:, bi])
            contrib = np.zeros(E)
            # (a charger without a vehicle reports last_charged_kwh = 0.0 (building.py:1378-1389), so the `no_car_charging` term never fires)
            soc_prev, soc_now, cap, mincap = info['soc_prev'][:, k], info['soc_now'][:, k], info['capacity'][:, k], info['min_capacity'][:, k]
            # np.float32 soc * python capacity + python kWh is float32 arithmetic (python-float initial_soc at t == 0: float64)
            cur = (w32(soc_prev) * w32(cap) + w32(kwh)).astype(f64) if t > 0 else soc_prev * cap + kwh
            c_con = np.where((cur > cap) | (cur < mincap), w['battery_limits'] * mult, 0.0)
            req, hrs = info['required'][:, k], info['hours'][:, k]
            diff = w32(soc_now).astype(f64) - req
            diff_kwh = diff * cap
            max_c, max_d = q[CP['MAX_C']] * hrs, q[CP['MAX_D']] * hrs
            c_con = c_con + np.where(diff_kwh > max_c, w['soc_impossible'] * mult, 0.0)
            at_dep = hrs == 0
            c_con = c_con + np.where(at_dep & (-0.25 < diff) & (diff <= -0.10), 2 * w['soc_under'] * mult, 0.0)
            c_con = c_con + np.where(at_dep & (diff <= -0.25), (w['soc_under'] ** 2) * mult, 0.0)
            c_con = c_con + np.where(at_dep & (-0.10 < diff) & (diff <= 0.10), w['close_soc'] * mult, 0.0)
            with np.errstate(divide='ignore'):
                c_con = c_con + np.where(np.abs(diff_kwh) <= np.maximum(max_c, max_d), w['close_soc'] * mult * (1.0 / (hrs + 0.1)), 0.0)
            c_con = c_con + np.where((kwh > 0) & (net < 0), w['extra_self_production'] * mult, 0.0)
            c_con = c_con + np.where((kwh < 0) & (net < 0), -0.5 * w['extra_self_production'] * mult, 0.0)
            c_con = c_con + np.where((kwh < 0) & (net > 0), w['self_ev_consumption'] * mult, 0.0)
            c_con = c_con + np.where((kwh > 0) & (net > 0), -0.5 * w['self_ev_consumption'] * mult, 0.0)
            out[:, bi] = out[:, bi] + contrib + np.where(con, c_con, 0.0)
        # charging-constraint penalty (reward_function.py:431-434): the violation the building reports, times the coefficient
        coef = self.attrs.get('charging_constraint_penalty_coefficient')
        coef = 1.0 if coef is None else float(coef)
        for k, bi in enumerate(env.evd.get('cc_building', ())):
            if env.spec.buildings[bi].charging_constraints.expose_violation:
                viol = env.cc_state[:, k * S.CC_SLOTS + S.CC_SLOTS - 1]
                out[:, bi] = out[:, bi] - np.where(viol > 0.0, viol * coef, 0.0)
        if env.central:
            tot = np.zeros((E, 1))
            for bi in range(B):
                tot[:, 0] = tot[:, 0] + out[:, bi]
            return tot
        return out

    def calculate(self, env, t, dyn, district):
        e = dyn[..., DYN['net_electricity_consumption']]
        k = self.kind
        if k == 'RewardFunction':
            ex = 1.0 if self.attrs.get('exponent') is None else self.attrs['exponent']
            r = -np.maximum(w32(e), np.float32(0.0)) ** np.float32(ex)          # np.float32 ** python float -> float32
        elif k == 'MARL':
            dsum = district[:, 0:1]            # python sum() of np.float32 in building order == the district total
            be = e * -1
            r = np.sign(be) * 0.01 * be ** 2 * np.fmax(0.0, dsum)               # np.array(..., dtype=float): float64
        elif k == 'IndependentSACReward':
            r = np.minimum(w32(e) * -1 ** 3, np.float32(0.0))   # operator precedence kept: -1**3 == -1 (reward_function.py:161)
        elif k == 'SolarPenaltyReward':
            r = self._solar_penalty(env, dyn)
        elif k == 'ComfortReward':
            r = self._comfort(env, t, dyn)
        elif k == 'SolarPenaltyAndComfortReward':
            co = self.attrs.get('coefficients') or [1.0, 1.0]
            r = self._solar_penalty(env, dyn).astype(f64) * co[0] + self._comfort(env, t, dyn).astype(f64) * co[1]
        elif k == 'Electric_Vehicles_Reward_Function':
            return self._ev_reward(env, t, dyn, district)
        else:
            raise NotImplementedError(k)
        if env.central:
            out = np.zeros((r.shape[0], 1), dtype=r.dtype)
            for bi in range(r.shape[1]):       # python sum() in building order, in the reward's own dtype
                out[:, 0] = out[:, 0] + r[:, bi]
            return out
        return r


class OracleMultiReward:
    """MultiBuildingRewardFunction (reward_function.py:90-117, built at citylearn.py:2106-2141): building i is rewarded by its own
    function, each evaluated on that building's observations alone (so a district sum is that building's own value)."""

    def __init__(self, per_building):
        self.per_building = per_building          # [OracleReward] in building order

    def calculate(self, env, t, dyn, district):
        cols = []
        for bi, rf in enumerate(self.per_building):
            own = dyn[:, bi:bi + 1, :]
            own_district = np.stack([own[:, 0, DYN['net_electricity_consumption']]] * 3, axis=1)
            sub = _OneBuilding(env, bi)
            cols.append(np.asarray(rf.calculate(sub, t, own, own_district), dtype=np.float64)[:, 0])
        return np.stack(cols, axis=1).astype(np.float32)


class _OneBuilding:
    """View of an OracleEnv restricted to building `bi` (what a per-building reward function sees)."""

    def __init__(self, env, bi):
        self._env, self._bi = env, bi
        self.central = False

    def col(self, name, t):
        return self._env.col(name, t)[..., self._bi:self._bi + 1]

    def col32(self, name, t):
        return self._env.col32(name, t)[..., self._bi:self._bi + 1]

    def P(self, name):
        return self._env.P(name)[..., self._bi:self._bi + 1]


def resolve_reward(spec: S.DistrictSpec):
    rt = spec.reward_type
    if isinstance(rt, dict):                      # per-building reward functions, 'default' fallback (citylearn.py:2106-2141)
        attrs = spec.reward_attributes or {}
        default_type = rt.get('default') or (next(iter(rt.values())) if rt else None)
        default_attrs = attrs.get('default')
        if default_attrs is None and attrs:
            default_attrs = next(iter(attrs.values()))
        per = []
        for b in spec.buildings:
            r_type = rt.get(b.name, default_type)
            per.append(OracleReward(r_type.split('.')[-1], **(attrs.get(b.name, default_attrs) or {})))
        return OracleMultiReward(per)
    name = rt.split('.')[-1] if isinstance(rt, str) else getattr(rt, '__name__', str(rt))
    return OracleReward(name, **(spec.reward_attributes or {}))
