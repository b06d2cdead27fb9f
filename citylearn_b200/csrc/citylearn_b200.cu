// citylearn_b200.cu - sm_100a kernels + C ABI of the CityLearn step path (see include/citylearn_b200.h).
//
// Data layout in HBM (all per cl_env handle, i.e. per GPU shard of E envs):
//   table    float  [n_rows][Wp]      one row per dataset time step, Wp = W rounded up to 4 floats (16 B) so that the
//                                     rows of step t and t+1 are ONE contiguous, 16B-aligned span -> a single TMA bulk copy
//   params   float / double [CL_NPARAM][B]   parameter-major so that the B buildings of a warp read consecutive words
//   iparams  int32  [CL_NIPARAM][B]
//   obs_desc int4   [L]  (+ tcol int32 [L]: the same layout compiled to "template columns" for the fast path)
//   outage   float  [B][T]            power-outage signal of the running episode
//   start    int32  [E]               table row of time step 0 of every env
//   state    float  [6][E*B]  (+ double [2][E*B] in CL_PRECISION_FP64)   unit index u = e * B + b (building fastest)
//   lstm     float  [90][E*B]         h, c of both layers and the two fed-back input windows (LSTM dynamics districts)
//   ev_*     float / double / u8      vehicles' soc[t-1] / soc[t], degraded capacity, efficiency, flags; washing-machine flags;
//                                     charging-constraint headroom / violation [E][n_constrained * 6] (cl_ev_desc districts)
//   obs_tab  float  [n_rows][L]       complete (wrapper-transformed) reference-parity observation row of every time step, built
//                                     once; the step kernel only moves it (TMA load -> TMA stores)
//   kpi_*    double [E][B][8], [E][2][15]   optional online KPI accumulators (cl_kpi_*)
//
// Kernels: `advance_kernel` (K >= 1 consecutive time steps in one launch: cl_step is K = 1, cl_rollout any K), `reset_kernel`,
// and the small one-off / per-step helpers `build_obs_table_kernel`, `district_finish_kernel`, `kpi_accumulate_kernel`.
// Thread mapping: one thread per unit; a block owns `envs_per_block` consecutive envs x all B buildings, so its slice of
// actions [E][A], rewards [E][B], state [.][E*B] and observations [E][L] is one contiguous range each (coalesced), and the
// district sums of an env never leave the block (shared memory, summed in building order like the reference's sum()).
// Districts wider than one block (WIDE instantiation) split the buildings of an env into tiles, one CTA each: independent CTAs
// with deferred district sums, or a thread-block cluster exchanging partial sums through distributed shared memory when a
// reward reads the district sum inside the step.
// LSTM dynamics: `lstm_update_mma` (warp-level tensor-core cell: mma.sync m16n8k8, 3xTF32, operand fragments in shared memory) where a
// block's warps each sit on one building, else the scalar `lstm_update`.  Host steps: `cl_step_host` (one call per step; page-locked host
// memory accessed in place by the kernels).  Building-sharded districts: `exchange_sum` (district sums completed over NVLink peer memory).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "unit_physics.cuh"

namespace cl {

// ------------------------------------------------------------------------------------------------------------------
// device-side district description (passed by value to kernels)
// ------------------------------------------------------------------------------------------------------------------
struct Dev {
    int B, E, U, n_rows, W, Wp, A, L, T;
    int central, reward_id, stale, envs_per_block, uniform_start, start0, has_outage, any_dynamics, lstm_smem;
    int lstm_const;        // 1: this district's packed LSTM weights are in the constant bank (c_lstm_w) for this launch
    int curve_nmax;        // max number of points of any battery curve of the district (uniform bound of the segment search)
    const uint8_t* curve_lut;   // [B][2][kCurveLutStride] uniform-grid index of the curve abscissae or nullptr (unit_physics.cuh)
    // buildings usually share a handful of distinct battery curve sets (one, in the bundled datasets and the synthetic districts): with
    // n_curves > 0 shared memory holds the n_curves distinct tables only, curve_id[b] names building b's, curve_rep[c] a building that has it
    int n_curves;                    // 0: one table per building of the block / tile (no sharing)
    const int32_t* curve_id;         // [B]
    const int32_t* curve_rep;        // [n_curves]
    // building tiles: a district wider than one block is split into `tiles` tiles of `tile_b` buildings, one CTA per tile,
    // the CTAs of an env forming a thread-block cluster (tiles == 1: tile_b == B, Lt == L, no cluster)
    int tiles, tile_b, Lt;
    int coupled;           // wide districts, set per launch: 1 = a reward needs cross-tile sums in-step (cluster + DSMEM path);
                           // 0 = tiles are independent, per-tile partial district sums go to a scratch buffer (district_finish_kernel)
    int tab_layout;        // 1: observation rows come from obs_tab (no gather-column staging in shared memory)
    const int32_t* tile_k; // [tiles + 1] observation-row range of every tile (wide districts only)
    // optional wrapper semantics fused into the observation writers / action fetch (cl_set_transforms; nullptr: identity)
    const cl_obs_transform* obs_t;   // [L]
    const float* act_range;          // [A] normalised action a in [0, 1] -> a * range + low (wrappers.py:208-222)
    const float* act_low;            // [A]
    // reference-parity observation rows do not depend on the env or the actions: row r of `obs_tab` is the complete (transformed)
    // observation of the time step whose table row is r, built once (build_obs_table_kernel).  The helper warp then only moves
    // it: TMA load of the block's column range -> shared memory -> TMA stores into the envs' rows.
    const float* obs_tab;            // [n_rows][obs_pitch] or nullptr (gather path)
    int obs_pitch;                   // floats per row (L rounded up to 4)
    int n_out_cols;                  // observation columns that carry an outage signal (patched per step: they change per episode)
    const int32_t* out_cols;         // [n_out_cols] column index k
    // stale_observations = 0 with an observation table: every env of a block owns a shared-memory image of its row (TMA-loaded
    // from obs_tab), the physics threads patch their action-dependent (DYN) columns into it, one TMA bulk store per env row
    int fresh_slots;                 // 1: the shared-memory layout holds 2 x envs_per_block row images (else 2 x 1)
    const int2* dyn_cols;            // [dyn_off[B]] (observation column k, cl_dyn slot), grouped by building
    const int32_t* dyn_off;          // [B + 1]
    float rp[8];
    const float* table;
    const float* pf;       // [NPARAM][B]
    const double* pd;      // [NPARAM][B]
    const int32_t* ip;     // [NIPARAM][B]
    const int4* desc;      // [L]
    const int32_t* tcol;   // [L] >= 0: table column, -1: zero (stale DYN slot), <= -2: outage signal of building (-2 - tcol)
    const float* outage;   // [B][T] or nullptr
    const int32_t* start;  // [E]
    const int32_t* t_dev;  // [1] device-resident time step (cl_step_device: launches with t0 < 0 read it)
    // online KPI accumulators fused into the step (cl_kpi_enable on a district that is neither building-tiled nor LSTM-driven): the
    // running sums live in shared memory for the steps of a launch and are folded into these arrays when it ends
    double* kpi_unit;                // [E][B][CL_NKPI_UNIT] or nullptr
    double* kpi_env;                 // [E][2][CL_NKPI_ENV]
    int kpi_smem;                    // 1: the shared-memory layout carries the accumulators
    // electric vehicles / chargers / washing machines (cl_ev_desc; unit_physics.cuh: charger_step)
    int ev_n, ch_n, wm_n;
    const double* ev_pd;             // [ev_n][CL_NPARAM] vehicle battery parameters
    const int32_t* ev_ip;            // [ev_n][2] curve point counts
    const int32_t* ev_cols;          // [ev_n][4] table columns (association SOC, pre-connection SOC, -, episode-start SOC)
    const double* ev_drift;          // [n_rows][ev_n]
    const int32_t* ch_off;           // [B + 1] chargers of building b: [ch_off[b], ch_off[b + 1])
    const int32_t* ch_action;        // [ch_n]
    const int32_t* ch_cols;          // [ch_n][4]
    const double* ch_pd;             // [ch_n][CL_NCHP]
    const int32_t* wm_off;           // [B + 1]
    const int32_t* wm_action;        // [wm_n]
    const int32_t* wm_cols;          // [wm_n][4]
    float* ev_sf;                    // [2][E * ev_n] soc[t-1], soc[t] entries of every vehicle
    double* ev_sd;                   // [2][E * ev_n] degraded capacity, round-trip efficiency of the last charge
    uint8_t* ev_flag;                // [E * ev_n] 1: the vehicle's battery has charged before (its efficiency / capacity are np.float64 from then on)
    uint8_t* wm_flag;                // [E * wm_n] 1: a cycle was started in the current window
    // charging constraints (cl_ev_desc.cc_*)
    int cc_n, obs_state;             // obs_state: the observation row has CL_OBS_STATE entries (per-env values: no observation table)
    const int32_t* cc_index;         // [B] index of the building's constraint record or -1
    const double* cc_limits;         // [cc_n][1 + CL_MAX_PHASES]
    const int32_t* cc_members;       // [cc_n][CL_MAX_PHASES][CL_MAX_CHARGERS_PER_BUILDING]
    const int32_t* cc_flags;         // [cc_n]
    float* cc_state;                 // [E][cc_n * CL_CC_SLOTS]
    // building-sharded districts (cl_exchange_*): this handle owns SOME buildings of every env; the per-env district sums are completed
    // inside the step by an all-gather of the ranks' partial sums through peer memory (NVLink): every (quantity, env) value travels as
    // ONE 8-byte {value, epoch} store into every peer's slot array - data and flag in one NVLink transaction, no fence, no second
    // round trip - and the reader spins on the epoch.  Slot (parity p, source rank r, env e, quantity q) = ((p * n + r) * E + e) * 3 + q.
    int x_n, x_rank;                 // ranks sharing the district (0: not sharded), this handle's rank
    unsigned x_epoch;                // epoch of the launch's first step minus 1 (epochs count steps since cl_exchange_create, never repeat)
    uint2* const* x_peers;           // [x_n] device array: base of every rank's slot array as mapped on THIS device (x_peers[x_rank]: own)
    int32_t* x_err;                  // [1] spin time-outs (a peer that never arrives must not hang the GPU)
    const float* lstm_w;   // packed LSTM weights [B][kLstmStride] (buildings without dynamics: zeros)
    float* lst;            // LSTM state [kLstmStateFloats][U] (dynamics districts only)
    float* st;             // [6][U]
    double* dst;           // [2][U] (fp64 mode)
};

enum { ST_SOC_B = 0, ST_CAP_DEG = 1, ST_RTE_B = 2, ST_SOC_CS = 3, ST_SOC_HS = 4, ST_SOC_DS = 5, ST_N = 6 };

template <typename R> struct PSel;
template <> struct PSel<float> { static __device__ __forceinline__ const float* p(const Dev& d) { return d.pf; } };
template <> struct PSel<double> { static __device__ __forceinline__ const double* p(const Dev& d) { return d.pd; } };

// per-thread launch constants: the unit's building parameters, table columns and action slots (registers)
template <typename R> struct UnitCtx {
    BuildingParams<R> p;
    R pv;
    int c_nsl, c_solar, c_price, c_carbon, c_tin;
    int a_es;
    int c_dhw, c_cool, c_heat, c_tout, c_hvac;
    int a_cd, a_hd, a_coh, a_cs, a_hs, a_ds;
    int c_coolsp, c_heatsp, c_band;
    // LSTM dynamics
    int dyn_c_inputs, dyn_n_inputs, dyn_slot_tin, dyn_slot_cdem, dyn_lookback;
    float tin_min, tin_range, cdem_min, cdem_range;
};

template <typename R, bool THERMAL>
__device__ __forceinline__ void load_ctx(const Dev& d, int b, UnitCtx<R>& c) {
    const auto* P = PSel<R>::p(d);
    const int B = d.B;
    BuildingParams<R>& p = c.p;
#define LD(k) ((R)__ldg(P + (k) * B + b))
#define LDI(k) (__ldg(d.ip + (k) * B + b))
    p.bat_capacity = LD(CL_P_BAT_CAPACITY); p.bat_pnom = LD(CL_P_BAT_NOMINAL_POWER); p.bat_loss = LD(CL_P_BAT_LOSS);
    p.bat_clc = LD(CL_P_BAT_CLC); p.bat_dod = LD(CL_P_BAT_DOD);
    p.ratio = LD(CL_P_TIME_STEP_RATIO); p.hours = LD(CL_P_HOURS_PER_STEP);
    p.flags = LDI(CL_IP_FLAGS); p.pe_n = LDI(CL_IP_PE_N); p.cp_n = LDI(CL_IP_CP_N);
    c.pv = LD(CL_P_PV_NOMINAL_POWER);
    c.c_nsl = LDI(CL_IP_C_NSL); c.c_solar = LDI(CL_IP_C_SOLAR); c.c_price = LDI(CL_IP_C_PRICE); c.c_carbon = LDI(CL_IP_C_CARBON);
    c.c_tin = LDI(CL_IP_C_T_IN);
    c.a_es = LDI(CL_IP_A_ELECTRICAL_STORAGE);
    c.c_coolsp = LDI(CL_IP_C_COOL_SP); c.c_heatsp = LDI(CL_IP_C_HEAT_SP); c.c_band = LDI(CL_IP_C_COMFORT_BAND);
    c.c_hvac = LDI(CL_IP_C_HVAC_MODE);
    if (THERMAL) {
        p.cd_pnom = LD(CL_P_CD_NOMINAL_POWER); p.cd_cop_num = LD(CL_P_CD_COP_NUM); p.cd_target = LD(CL_P_CD_TARGET);
        p.hd_pnom = LD(CL_P_HD_NOMINAL_POWER); p.hd_cop_num = LD(CL_P_HD_COP_NUM); p.hd_target = LD(CL_P_HD_TARGET); p.hd_eff = LD(CL_P_HD_EFFICIENCY);
        p.dd_pnom = LD(CL_P_DD_NOMINAL_POWER); p.dd_cop_num = LD(CL_P_DD_COP_NUM); p.dd_target = LD(CL_P_DD_TARGET); p.dd_eff = LD(CL_P_DD_EFFICIENCY);
        p.cs.capacity = LD(CL_P_CS_CAPACITY); p.cs.efficiency = LD(CL_P_CS_EFFICIENCY); p.cs.loss = LD(CL_P_CS_LOSS);
        p.cs.max_in = LD(CL_P_CS_MAX_IN); p.cs.max_out = LD(CL_P_CS_MAX_OUT);
        p.cs.has_max_in = p.flags & CL_F_CS_HAS_MAX_IN; p.cs.has_max_out = p.flags & CL_F_CS_HAS_MAX_OUT;
        p.hs.capacity = LD(CL_P_HS_CAPACITY); p.hs.efficiency = LD(CL_P_HS_EFFICIENCY); p.hs.loss = LD(CL_P_HS_LOSS);
        p.hs.max_in = LD(CL_P_HS_MAX_IN); p.hs.max_out = LD(CL_P_HS_MAX_OUT);
        p.hs.has_max_in = p.flags & CL_F_HS_HAS_MAX_IN; p.hs.has_max_out = p.flags & CL_F_HS_HAS_MAX_OUT;
        p.ds.capacity = LD(CL_P_DS_CAPACITY); p.ds.efficiency = LD(CL_P_DS_EFFICIENCY); p.ds.loss = LD(CL_P_DS_LOSS);
        p.ds.max_in = LD(CL_P_DS_MAX_IN); p.ds.max_out = LD(CL_P_DS_MAX_OUT);
        p.ds.has_max_in = p.flags & CL_F_DS_HAS_MAX_IN; p.ds.has_max_out = p.flags & CL_F_DS_HAS_MAX_OUT;
        c.c_dhw = LDI(CL_IP_C_DHW_DEMAND); c.c_cool = LDI(CL_IP_C_COOLING_DEMAND); c.c_heat = LDI(CL_IP_C_HEATING_DEMAND);
        c.c_tout = LDI(CL_IP_C_T_OUT);
        c.a_cd = LDI(CL_IP_A_COOLING_DEVICE); c.a_hd = LDI(CL_IP_A_HEATING_DEVICE); c.a_coh = LDI(CL_IP_A_COOLING_OR_HEATING_DEVICE);
        c.a_cs = LDI(CL_IP_A_COOLING_STORAGE); c.a_hs = LDI(CL_IP_A_HEATING_STORAGE); c.a_ds = LDI(CL_IP_A_DHW_STORAGE);
        c.dyn_c_inputs = LDI(CL_IP_DYN_C_INPUTS); c.dyn_n_inputs = LDI(CL_IP_DYN_N_INPUTS); c.dyn_slot_tin = LDI(CL_IP_DYN_SLOT_TIN);
        c.dyn_slot_cdem = LDI(CL_IP_DYN_SLOT_CDEM); c.dyn_lookback = LDI(CL_IP_DYN_LOOKBACK);
        // normalisation constants: (float32 - python min) / (python max - python min) is float32 arithmetic (building.py:3070-3078)
        const double tmin = __ldg(d.pd + CL_P_DYN_TIN_MIN * B + b), tmax = __ldg(d.pd + CL_P_DYN_TIN_MAX * B + b);
        const double cmin = __ldg(d.pd + CL_P_DYN_CDEM_MIN * B + b), cmax = __ldg(d.pd + CL_P_DYN_CDEM_MAX * B + b);
        c.tin_min = (float)tmin; c.tin_range = (float)(tmax - tmin); c.cdem_min = (float)cmin; c.cdem_range = (float)(cmax - cmin);
    }
#undef LD
#undef LDI
    derive_params(p);
}

template <typename R, bool THERMAL>
__device__ __forceinline__ void load_state(const Dev& d, int u, UnitState<R>& s) {
    const int U = d.U;
    s.soc_b = (R)d.st[ST_SOC_B * U + u];
    if (sizeof(R) == 8) { s.cap_deg = (R)d.dst[u]; s.rte_b = (R)d.dst[U + u]; }
    else { s.cap_deg = (R)d.st[ST_CAP_DEG * U + u]; s.rte_b = (R)d.st[ST_RTE_B * U + u]; }
    if (THERMAL) { s.soc_cs = (R)d.st[ST_SOC_CS * U + u]; s.soc_hs = (R)d.st[ST_SOC_HS * U + u]; s.soc_ds = (R)d.st[ST_SOC_DS * U + u]; }
    else { s.soc_cs = s.soc_hs = s.soc_ds = (R)0; }
}

template <typename R, bool THERMAL>
__device__ __forceinline__ void store_state(const Dev& d, int u, const UnitState<R>& s) {
    const int U = d.U;
    d.st[ST_SOC_B * U + u] = (float)s.soc_b;
    if (sizeof(R) == 8) { d.dst[u] = (double)s.cap_deg; d.dst[U + u] = (double)s.rte_b; }
    else { d.st[ST_CAP_DEG * U + u] = (float)s.cap_deg; d.st[ST_RTE_B * U + u] = (float)s.rte_b; }
    if (THERMAL) { d.st[ST_SOC_CS * U + u] = (float)s.soc_cs; d.st[ST_SOC_HS * U + u] = (float)s.soc_hs; d.st[ST_SOC_DS * U + u] = (float)s.soc_ds; }
}

// a time row: TMA-staged in shared memory (lock-step episode windows; `s` is its shared-window address - explicit ld.shared instead
// of a generic load, which is tracked like a global one) or straight from the table in global memory (`s` == 0)
struct RowRef {
    const float* g; uint32_t s;
    __device__ __forceinline__ float operator[](int col) const {
        float v;
        if (s != 0u) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(s + 4u * (uint32_t)col));   // volatile: stays behind the row's mbarrier wait
        else v = __ldg(g + col);
        return v;
    }
};

// exogenous inputs of a unit at time step t; `row` points at the time row (shared or global memory)
template <typename R, bool THERMAL>
__device__ __forceinline__ void load_inputs(const Dev& d, const UnitCtx<R>& c, const RowRef row, int b, int t, UnitInputs<R>& in,
                                            bool solar_from_block = false) {
    in.nsl = (R)row[c.c_nsl];
    // building.py:2554, energy_model.py:488; with lock-step rows the helper warp computes it once per building
    in.solar = solar_from_block ? (R)0 : -dvd(c.pv * (R)row[c.c_solar], (R)1000);
    in.price = (R)row[c.c_price];
    in.carbon = (R)row[c.c_carbon];
    if (THERMAL) {
        in.dhw_demand = (R)row[c.c_dhw]; in.cooling_demand = (R)row[c.c_cool]; in.heating_demand = (R)row[c.c_heat];
        in.t_out = (R)row[c.c_tout]; in.hvac_mode = (int32_t)row[c.c_hvac];
    } else {
        in.dhw_demand = in.cooling_demand = in.heating_demand = (R)0; in.t_out = (R)0; in.hvac_mode = 0;
    }
    in.outage = d.has_outage && (c.p.flags & CL_F_SIMULATE_OUTAGE) && __ldg(d.outage + b * d.T + t) > 0.f;
    in.control_cooling_demand = false;
    in.control_heating_demand = false;
}

// raw action values of one unit for one step (registers); fetched ONE STEP AHEAD so that the HBM latency of the read hides
// behind the physics of the current step
struct RawActions { float es, cd, hd, coh, cs, hs, ds; };

template <typename R, bool THERMAL>
__device__ __forceinline__ void fetch_actions(const Dev& d, const UnitCtx<R>& c, const float* act_row, RawActions& a) {
    // loads only: nothing here may consume the loaded value (even a predicated-off instruction waits for its operands, which would
    // turn the one-step-ahead prefetch into a ~600-cycle stall); the action transform is applied where the value is used
    // `asm volatile`: the compiler may otherwise sink the load to its use one step later (it did: the action load sat at the top of the
    // next iteration with a full global-memory latency exposed - ncu: 10 % of the stall samples on the F2F that consumes it)
    auto get = [&](int col) -> float {
        float v = 0.f;
        if (col >= 0) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(act_row + col));
        return v;
    };
    a.es = get(c.a_es);
    if (THERMAL) {
        a.cd = get(c.a_cd); a.hd = get(c.a_hd); a.coh = get(c.a_coh);
        a.cs = get(c.a_cs); a.hs = get(c.a_hs); a.ds = get(c.a_ds);
    }
}

// observation transform of column k (cl_obs_transform): periodic sin / cos, affine min-max, clip - NaN passes through
__device__ __forceinline__ float transform_obs(const cl_obs_transform* t, int k, float v) {
    const cl_obs_transform* x = t + k;
    const int fn = __ldg(&x->fn);
    if (fn != CL_OBS_FN_IDENTITY) {
        // periodic observations are 1 .. x_max, i.e. the angle is in (0, 2 pi]: fold it into (-pi, pi] and use the hardware
        // approximations (absolute error < 5e-7 there) - the library sinf / cosf would drag their slow-path local array and a
        // lower register budget into every instantiation of the step kernel
        float a = v * __ldg(&x->w);
        if (a > 3.14159265358979f) a -= 6.28318530717959f;
        v = fn == CL_OBS_FN_SIN ? __sinf(a) : __cosf(a);
    }
    v = v * __ldg(&x->scale) + __ldg(&x->offset);
    const float lo = __ldg(&x->lo), hi = __ldg(&x->hi);
    return v < lo ? lo : (v > hi ? hi : v);
}

// inactive storage actions are 0, inactive device actions NaN (building.py:1555-1564)
template <typename R, bool THERMAL>
__device__ __forceinline__ void apply_actions(const Dev& d, const UnitCtx<R>& c, RawActions a, UnitInputs<R>& in) {
    if (d.act_range != nullptr) {
        // with an action transform the caller's values are fractions of the action range (NormalizedActionWrapper)
        auto tr = [&](int col, float& v) { if (col >= 0) v = v * __ldg(d.act_range + col) + __ldg(d.act_low + col); };
        tr(c.a_es, a.es);
        if (THERMAL) { tr(c.a_cd, a.cd); tr(c.a_hd, a.hd); tr(c.a_coh, a.coh); tr(c.a_cs, a.cs); tr(c.a_hs, a.hs); tr(c.a_ds, a.ds); }
    }
    in.a_es = (R)a.es;
    in.a_cooling_device = in.a_heating_device = (R)NAN;
    in.a_cs = in.a_hs = in.a_ds = (R)0;
    if (THERMAL) {
        if (c.a_cd >= 0) in.a_cooling_device = (R)a.cd;
        if (c.a_hd >= 0) in.a_heating_device = (R)a.hd;
        if (c.a_coh >= 0) {   // building.py:1550-1553
            const R v = (R)a.coh;
            in.a_cooling_device = fabs(rmin(v, (R)0));
            in.a_heating_device = fabs(rmax(v, (R)0));
        }
        in.a_cs = (R)a.cs; in.a_hs = (R)a.hs; in.a_ds = (R)a.ds;
    }
}

// value of cl_dyn slot `slot` of a unit from a step / time-0 result
template <typename R>
__device__ __forceinline__ float dyn_value(int slot, const BuildingParams<R>& p, const UnitState<R>& s, const UnitResult<R>& o, R t_in) {
    using N = Num<R>;
    switch (slot) {
        case CL_DYN_ELECTRICAL_STORAGE_SOC: return (float)s.soc_b;
        case CL_DYN_COOLING_STORAGE_SOC: return (float)s.soc_cs;
        case CL_DYN_HEATING_STORAGE_SOC: return (float)s.soc_hs;
        case CL_DYN_DHW_STORAGE_SOC: return (float)s.soc_ds;
        case CL_DYN_NET_ELECTRICITY_CONSUMPTION: return (float)o.net;
        case CL_DYN_COOLING_DEMAND: return (float)(o.e_from_cool + fabs(rmin(o.eb_cs, (R)0)));
        case CL_DYN_HEATING_DEMAND: return (float)(o.e_from_heat + fabs(rmin(o.eb_hs, (R)0)));
        case CL_DYN_DHW_DEMAND: return (float)(o.e_from_dhw + fabs(rmin(o.eb_ds, (R)0)));
        case CL_DYN_COOLING_ELECTRICITY_CONSUMPTION: return (float)(o.ec_cool * p.ratio);
        case CL_DYN_HEATING_ELECTRICITY_CONSUMPTION: return (float)(o.ec_heat * p.ratio);
        case CL_DYN_DHW_ELECTRICITY_CONSUMPTION: return (float)(o.ec_dhw * p.ratio);
        case CL_DYN_COOLING_STORAGE_ELECTRICITY_CONSUMPTION: return (float)N::div32(o.eb_cs, o.eff_cool);
        case CL_DYN_HEATING_STORAGE_ELECTRICITY_CONSUMPTION: return (float)N::div32(o.eb_hs, o.eff_heat);
        case CL_DYN_DHW_STORAGE_ELECTRICITY_CONSUMPTION: return (float)N::div32(o.eb_ds, o.eff_dhw);
        case CL_DYN_ELECTRICAL_STORAGE_ELECTRICITY_CONSUMPTION: return (float)(o.ec_bat * p.ratio);
        case CL_DYN_INDOOR_DRY_BULB_TEMPERATURE: return (float)t_in;
        case CL_DYN_NON_SHIFTABLE_LOAD_ELECTRICITY_CONSUMPTION: return (float)(o.ec_nsl * p.ratio);
        case CL_DYN_ELECTRICAL_STORAGE_ENERGY_BALANCE: return (float)o.eb_bat;
        case CL_DYN_COOLING_STORAGE_ENERGY_BALANCE: return (float)o.eb_cs;
        case CL_DYN_HEATING_STORAGE_ENERGY_BALANCE: return (float)o.eb_hs;
        case CL_DYN_DHW_STORAGE_ENERGY_BALANCE: return (float)o.eb_ds;
        case CL_DYN_NET_ELECTRICITY_CONSUMPTION_COST: return (float)o.cost;
        case CL_DYN_NET_ELECTRICITY_CONSUMPTION_EMISSION: return (float)o.emission;
        case CL_DYN_ELECTRICAL_STORAGE_DEGRADED_CAPACITY: return (float)s.cap_deg;
        case CL_DYN_ENERGY_TO_NON_SHIFTABLE_LOAD: return (float)o.e_to_nsl;
        case CL_DYN_COOLING_DEMAND_SERIES: return (float)o.cool_dem;
        case CL_DYN_HEATING_DEMAND_SERIES: return (float)o.heat_dem;
        default: return 0.f;
    }
}
// all dyn values of a unit (cl_dyn order)
template <typename R>
__device__ __forceinline__ void fill_dyn(const BuildingParams<R>& p, const UnitState<R>& s, const UnitResult<R>& o, R t_in, float* dyn) {
#pragma unroll
    for (int j = 0; j < CL_NDYN; ++j) dyn[j] = dyn_value<R>(j, p, s, o, t_in);
}

// ------------------------------------------------------------------------------------------------------------------
// reward functions (citylearn/reward_function.py); observation values are np.float32 there -> float arithmetic,
// except MARL which converts to float64 (np.array(..., dtype=float)).
// ------------------------------------------------------------------------------------------------------------------
struct RewardIn {
    float net, soc_b, soc_cs, soc_hs, soc_ds, cool_dem, heat_dem, t_in, cool_sp, heat_sp, band_series;
    float cap_b, cap_cs, cap_hs, cap_ds;
    int hvac_mode;
    float district_net;
};

__device__ __forceinline__ float comfort_reward(const RewardIn& r, const float* rp) {
    const float band = isnan(rp[0]) ? r.band_series : rp[0];
    const float lo_e = rp[1], hi_e = rp[2];
    const bool heating = r.heat_dem > r.cool_dem;
    const float T = r.t_in;
    if (r.hvac_mode == 1 || r.hvac_mode == 2) {
        const float sp = r.hvac_mode == 1 ? r.cool_sp : r.heat_sp;
        const float lower = sp - band, upper = sp + band;
        const float delta = fabsf(T - sp);
        if (T < lower) return -powf(delta, r.hvac_mode == 2 ? lo_e : hi_e);
        if (lower <= T && T < sp) return heating ? 0.f : -delta;
        if (sp <= T && T <= upper) return heating ? -delta : 0.f;
        return -powf(delta, heating ? hi_e : lo_e);
    }
    const float lower = r.heat_sp - band, upper = r.cool_sp + band;
    const float cd = T - r.cool_sp, hd = T - r.heat_sp;
    if (T < lower) return -powf(fabsf(hd), !heating ? hi_e : lo_e);
    if (lower <= T && T < r.heat_sp) return -fabsf(hd);
    if (r.heat_sp <= T && T <= r.cool_sp) return 0.f;
    if (r.cool_sp < T && T < upper) return -fabsf(cd);
    return -powf(fabsf(cd), heating ? hi_e : lo_e);
}

__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : (x == 0.f ? 0.f : x)); }

__device__ __forceinline__ float solar_penalty_reward(const RewardIn& r) {
    const float e = r.net;
    float reward = 0.f;
    if (r.cap_cs > (float)kEps) reward += -(1.0f + sgnf(e) * r.soc_cs) * fabsf(e);
    if (r.cap_hs > (float)kEps) reward += -(1.0f + sgnf(e) * r.soc_hs) * fabsf(e);
    if (r.cap_ds > (float)kEps) reward += -(1.0f + sgnf(e) * r.soc_ds) * fabsf(e);
    if (r.cap_b > (float)kEps) reward += -(1.0f + sgnf(e) * r.soc_b) * fabsf(e);
    return reward;
}

// Electric_Vehicles_Reward_Function (citylearn/reward_function.py:389-523): what `electric_vehicles_chargers_dict` holds per charger at t
struct ChargerInfo { bool conn; float kwh, soc_now; double soc_prev, cap, min_cap, req, hrs, max_c, max_d; };
__device__ __forceinline__ float ev_reward(const ChargerInfo* ch, int n, float net, float district_net, int t, double penalty) {
    if (n == 0) return (float)(0.0 - penalty);                          // a building without chargers is rewarded 0 (:421-422)
    const double be = -(double)net;
    const double sg = be > 0 ? 1.0 : (be < 0 ? -1.0 : 0.0);
    const double marl = sg * 0.01 * (be * be) * fmax(0.0, (double)district_net);
    const double mult = 1.0 / (1.0 + fabs(marl));
    const double w_limits = -2.0, w_impossible = -10.0, w_under = -5.0, w_close = 10.0, w_self = 5.0, w_extra = 5.0;
    double total = 0.0;
    for (int i = 0; i < n; ++i) {
        const ChargerInfo& c = ch[i];
        if (!c.conn) continue;                                          // (its last_charged_kwh reads 0.0: `no_car_charging` never fires)
        double s = 0.0;
        const double kwh = (double)c.kwh;
        // soc[t-1] * capacity + kWh: float32 arithmetic on the np.float32 entry, float64 on the python-float initial SOC at t == 0
        const double cur = t > 0 ? (double)((float)c.soc_prev * (float)c.cap + c.kwh) : c.soc_prev * c.cap + kwh;
        if (cur > c.cap || cur < c.min_cap) s += w_limits * mult;
        const double diff = (double)c.soc_now - c.req, diff_kwh = diff * c.cap;
        const double max_c = c.max_c * c.hrs, max_d = c.max_d * c.hrs;
        if (diff_kwh > max_c) s += w_impossible * mult;
        if (c.hrs == 0.0) {
            if (-0.25 < diff && diff <= -0.10) s += 2 * w_under * mult;
            else if (diff <= -0.25) s += (w_under * w_under) * mult;
            else if (-0.10 < diff && diff <= 0.10) s += w_close * mult;
        }
        if (fabs(diff_kwh) <= fmax(max_c, max_d)) s += w_close * mult * (1.0 / (c.hrs + 0.1));
        if (kwh > 0 && net < 0.f) s += w_extra * mult; else if (kwh < 0 && net < 0.f) s += -0.5 * w_extra * mult;
        if (kwh < 0 && net > 0.f) s += w_self * mult; else if (kwh > 0 && net > 0.f) s += -0.5 * w_self * mult;
        total += s;
    }
    return (float)(total - penalty);                                    // charging-constraint violation x coefficient (:431-434)
}

__device__ __forceinline__ float unit_reward(int reward_id, const float* rp, const RewardIn& r) {
    switch (reward_id) {
        case CL_REWARD_DEFAULT: {
            const float m = fmaxf(r.net, 0.f);
            return rp[0] == 1.0f ? -m : -powf(m, rp[0]);
        }
        case CL_REWARD_MARL: {
            const double be = -(double)r.net;
            const double sg = be > 0 ? 1.0 : (be < 0 ? -1.0 : 0.0);
            return (float)(sg * 0.01 * (be * be) * fmax(0.0, (double)r.district_net));
        }
        case CL_REWARD_INDEPENDENT_SAC: return fminf(-r.net, 0.f);   // v * -1 ** 3 == -v (reward_function.py:161)
        case CL_REWARD_SOLAR_PENALTY: return solar_penalty_reward(r);
        case CL_REWARD_COMFORT: return comfort_reward(r, rp);
        case CL_REWARD_SOLAR_PENALTY_AND_COMFORT:
            return (float)((double)solar_penalty_reward(r) * (double)rp[3] + (double)comfort_reward(r, rp) * (double)rp[4]);
        default: return 0.f;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk) staging of the time rows t and t+1 into shared memory
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// thread-block cluster primitives (wide districts): rank, all-thread barrier, distributed-shared-memory loads
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_map(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank)); return r;
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t addr) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ double ld_cluster_f64(uint32_t addr) { double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory"); return v; }

// peer-memory exchange primitives (building-sharded districts): 8-byte {value, epoch} slots, written and polled with .volatile
// accesses (system-coherent L2 path; a 64-bit aligned store is one NVLink transaction - the flag can never be seen without its value)
__device__ __forceinline__ void st_slot(uint2* p, float v, unsigned epoch) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint2 ld_slot(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
// this rank's partial `acc` of (env e, quantity q) at step epoch `ep` -> the sum over all ranks, added in rank order (the same value
// on every rank).  A peer that does not deliver within ~2 s is counted in x_err and treated as 0.
__device__ __forceinline__ float exchange_sum(const Dev& d, float acc, int e, int q, unsigned ep) {
    const int n = d.x_n;
    const unsigned p = ep & 1u;
    const size_t mine = ((size_t)(p * n + d.x_rank) * d.E + e) * 3 + q;
    for (int r = 0; r < n; ++r) st_slot(d.x_peers[r] + mine, acc, ep);              // push to every rank (own copy included)
    const uint2* own = d.x_peers[d.x_rank];
    float tot = 0.f;
    for (int r = 0; r < n; ++r) {
        const uint2* src = own + ((size_t)(p * n + r) * d.E + e) * 3 + q;
        uint2 v = ld_slot(src);
        if (v.y != ep) {
            // a peer that has gone away must not hang the GPU: the first wait that runs out (~2 s) raises x_err, after which no wait
            // spins any more (the results are garbage from then on; cl_exchange_status reports it)
            const long long t_start = clock64();
            while ((v = ld_slot(src)).y != ep) {
                if (clock64() - t_start > 4000000000LL || *reinterpret_cast<volatile int32_t*>(d.x_err) != 0) { atomicAdd(d.x_err, 1); v.x = 0u; break; }
            }
        }
        tot += __uint_as_float(v.x);
    }
    return tot;
}

// ------------------------------------------------------------------------------------------------------------------
// observation writers: the rows of a block's envs are one contiguous span obs[e0*L .. (e0+n)*L)
// ------------------------------------------------------------------------------------------------------------------
// general path: any descriptor kind, per-env start rows, DYN values from shared memory (or zero)
__device__ __forceinline__ void write_obs_general(const Dev& d, float* obs, int e0, int n_env, int t_obs,
                                                   const float* dynbuf /* [n_env*nb][CL_NDYN] or nullptr */, int tid, int nt,
                                                   int k0, int k1, int b0, int nb) {
    const int L = d.L, Lt = k1 - k0;               // this block writes columns [k0, k1) of its envs' rows (whole rows: k0 = 0, k1 = L)
    const int total = n_env * Lt;
    int j = tid;
    int e_l = j / Lt, k = j - e_l * Lt;
    const int de = nt / Lt, dk = nt - de * Lt;
    for (; j < total; j += nt) {
        const int4 ds = __ldg(d.desc + k0 + k);
        float v;
        if (ds.x == CL_OBS_TS) {
            const int row = __ldg(d.start + e0 + e_l) + t_obs;
            v = __ldg(d.table + (size_t)row * d.Wp + ((t_obs == 0 && ds.z > 0) ? ds.z - 1 : ds.y));   // (b: the column to read on an episode's first row)
        } else if (ds.x == CL_OBS_DYN) {
            v = dynbuf ? dynbuf[(e_l * nb + (ds.w - b0)) * CL_NDYN + ds.y] : 0.f;
        } else if (ds.x == CL_OBS_STATE) {
            v = d.cc_state[(size_t)(e0 + e_l) * (d.cc_n * CL_CC_SLOTS) + ds.y];
        } else {
            v = d.has_outage ? __ldg(d.outage + ds.w * d.T + t_obs) : 0.f;
        }
        if (d.obs_t) v = transform_obs(d.obs_t, k0 + k, v);
        obs[(size_t)(e0 + e_l) * L + k0 + k] = v;
        e_l += de; k += dk;
        if (k >= Lt) { k -= Lt; e_l += 1; }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// shared memory carve-up (dynamic)
//   [mbarriers 64 B][curves B*kCurveTab (R)][bsolar 2*B (R)][rows 3*Wp][tcol Lp (int)][tmpl 2*Lp][red 2*3*nt][rsum nt][dsum 2*epb][dynbuf nt*NDYN (opt)]
// red / dsum / bsolar are double-buffered by step parity so that a step needs ONE block barrier (see advance_kernel).
// ------------------------------------------------------------------------------------------------------------------
// offsets in floats from the start of the dynamic shared memory (kept as plain ints so that every access is derived
// directly from the `extern __shared__` array and compiles to LDS/STS with 32-bit addressing, not generic loads)
struct SmemLayout {
    int curves, clut, bsolar, rows, tcol, tmpl, red, rsum, dsum, wpart, rpart, lstm, lstm_pre, lstm_frag, dynbuf, kpi_acc, kpi_nws, kpi_env, end, Lp;
};
__host__ __device__ __forceinline__ SmemLayout smem_layout(int B, int Wp, int L, int epb, int nt, int rsize, int lstm_smem = 0, int tab_layout = 0, int fresh_slots = 0, int n_curves = 0, int kpi = 0) {
    SmemLayout o;
    o.Lp = (L + 3) & ~3;
    int f = 32;                                  // 128 bytes of mbarriers: 3 time-row slots + 2 observation-row buffers + 6 hand-offs of the DEC instantiation
    const int nc = n_curves > 0 ? n_curves : B;       // (districts with vehicles pass B + ev_n: the vehicles' battery curves follow the buildings')
    o.curves = f; f += nc * kCurveTab * (rsize / 4);  // first: keeps doubles 8-byte aligned
    o.clut = f; f += nc * kCurveLutFloats;            // uniform-grid index of the curve abscissae (bytes)
    o.bsolar = f; f += ((2 * B * (rsize / 4)) + 3) & ~3;
    o.rows = f; f += 3 * Wp;
    o.tcol = f; f += tab_layout ? 0 : o.Lp;      // gather columns: not needed when the rows come from the observation table
    o.tmpl = f; f += 2 * o.Lp * (fresh_slots ? epb : 1);   // observation row image(s) of the step, source of the TMA bulk stores (double-buffered)
    o.red = f; f += 6 * nt;
    o.rsum = f; f += nt;
    o.dsum = f; f += (2 * epb + 3) & ~3;
    o.wpart = f; f += 2 * 3 * 32;                // wide districts: per-warp partial district sums [2][3][32 warps]
    o.rpart = f; f += 2 * 32 * 2;                // wide districts, central agent: per-warp partial reward sums [2][32] doubles (8-byte aligned: f is even)
    o.lstm = f; f += lstm_smem ? B * kLstmStride : 0;      // kLstmStride is a multiple of 4 floats: 16-byte aligned rows
    o.lstm_pre = f; f += lstm_smem ? B * kLstmPreRing * 64 : 0;   // per-building ring of shared layer-0 input projections
    o.lstm_frag = f; f += lstm_smem == 2 ? B * kLstmFragFloats : 0;   // tensor-core operand fragments of the recurrent / layer-1 matrices (lstm_update_mma)
    // with per-env row images the general writer's dynbuf is never live at the same time: it aliases them (cl_create checks the size)
    o.dynbuf = fresh_slots ? o.tmpl : f;
    // fused KPI accumulators (doubles; f is kept even): [CL_NKPI_UNIT][nt] running sums, [2][nt] baseline net of the step (by step
    // parity), [epb][2][CL_NKPI_ENV] district series state.  The general writer's dynbuf (appended after `end`) never coexists with them.
    f = (f + 1) & ~1;
    o.kpi_acc = f; f += kpi ? 2 * CL_NKPI_UNIT * nt : 0;
    o.kpi_nws = f; f += kpi ? 2 * 2 * nt : 0;
    o.kpi_env = f; f += kpi ? 2 * 2 * CL_NKPI_ENV * epb : 0;
    o.end = f;
    return o;
}
static size_t smem_bytes(const Dev& d, int nt, bool with_dyn, int rsize) {
    const SmemLayout o = smem_layout(d.tile_b, d.Wp, d.Lt, d.envs_per_block, nt, rsize, d.lstm_smem, d.tab_layout, d.fresh_slots, d.ev_n > 0 ? d.tile_b + d.ev_n : d.n_curves, d.kpi_smem);
    size_t n = sizeof(float) * (size_t)o.end;
    if (with_dyn && !d.fresh_slots) n += sizeof(float) * (size_t)nt * CL_NDYN;
    return n;
}

template <typename R, bool THERMAL>
__device__ __forceinline__ void reward_inputs(const Dev& d, const UnitCtx<R>& c, const UnitState<R>& s, const UnitResult<R>& o,
                                              const RowRef row, float t_in, RewardIn& ri) {
    const BuildingParams<R>& p = c.p;
    ri.net = (float)o.net; ri.district_net = 0.f;
    ri.soc_b = (float)s.soc_b; ri.soc_cs = (float)s.soc_cs; ri.soc_hs = (float)s.soc_hs; ri.soc_ds = (float)s.soc_ds;
    ri.cap_b = (float)p.bat_capacity;
    if (THERMAL) {
        ri.cap_cs = (float)p.cs.capacity; ri.cap_hs = (float)p.hs.capacity; ri.cap_ds = (float)p.ds.capacity;
        ri.cool_dem = (float)(o.e_from_cool + fabs(rmin(o.eb_cs, (R)0)));
        ri.heat_dem = (float)(o.e_from_heat + fabs(rmin(o.eb_hs, (R)0)));
    } else { ri.cap_cs = ri.cap_hs = ri.cap_ds = 0.f; ri.cool_dem = ri.heat_dem = 0.f; }
    ri.t_in = t_in;
    if (d.reward_id == CL_REWARD_COMFORT || d.reward_id == CL_REWARD_SOLAR_PENALTY_AND_COMFORT) {
        ri.cool_sp = row[c.c_coolsp]; ri.heat_sp = row[c.c_heatsp]; ri.band_series = row[c.c_band];
        ri.hvac_mode = (int)row[c.c_hvac];
    } else { ri.cool_sp = ri.heat_sp = ri.band_series = 0.f; ri.hvac_mode = 0; }
}

// ------------------------------------------------------------------------------------------------------------------
// LSTM dynamics of one unit after its physics at step t (building.py:2935-2942, 3000-3078).
// The two action-dependent inputs (cooling demand, indoor temperature) live in per-unit ring windows indexed by time step;
// the exogenous inputs come pre-normalised from the table.  Returns the indoor temperature of step t (prediction once the
// window is full, else the dataset value).
// ------------------------------------------------------------------------------------------------------------------
template <typename R, bool CONSTW>
__device__ __forceinline__ float lstm_update(const Dev& d, const UnitCtx<R>& c, const float* W, int u, int t, int row0 /* table row of step 0 */,
                                             float obs_cool_dem, float t_in_dataset, const float* pre /* this building's projection ring or nullptr */,
                                             int bslot /* building index: slot in c_lstm_w */) {
    const int U = d.U;
    const int L = c.dyn_lookback, ring = L + 1;
    const bool smem_w = d.lstm_smem != 0;                 // W then points into shared memory
    const uint32_t ws = smem_w ? smem_u32(W) : 0u;
    float* lst = d.lst;
    float* win_c = lst + (size_t)(4 * kLstmH) * U + u;                     // [ring][U]
    float* win_t = lst + (size_t)(4 * kLstmH + kLstmMaxLookback + 1) * U + u;
    // _update_dynamics_input: append the normalised observation of step t (float32 arithmetic)
    if (c.dyn_slot_cdem >= 0) win_c[(size_t)(t % ring) * U] = (obs_cool_dem - c.cdem_min) / c.cdem_range;
    win_t[(size_t)(t % ring) * U] = (t_in_dataset - c.tin_min) / c.tin_range;
    if (t < L) return t_in_dataset;                                         // window not full yet (building.py:2996-2998)
    float h0[kLstmH], h1[kLstmH], c0[kLstmH], c1[kLstmH];
#pragma unroll
    for (int j = 0; j < kLstmH; ++j) {
        h0[j] = lst[(size_t)j * U + u]; h1[j] = lst[(size_t)(kLstmH + j) * U + u];
        c0[j] = lst[(size_t)(2 * kLstmH + j) * U + u]; c1[j] = lst[(size_t)(3 * kLstmH + j) * U + u];
    }
    if (CONSTW && pre != nullptr && d.lstm_const) {
        // weights as immediate constant-bank operands: one fully unrolled instantiation per building slot (unit_physics.cuh)
        const uint32_t ps0 = smem_u32(pre);
        auto run = [&](auto BI) {
          if constexpr (CONSTW) {
#pragma unroll 1
            for (int sidx = 0; sidx < L; ++sidx) {
                const int tau = t - (L - 1) + sidx;
                const float xc = c.dyn_slot_cdem >= 0 ? win_c[(size_t)(tau % ring) * U] : 0.f;
                const float xt = win_t[(size_t)((tau - 1) % ring) * U];
                lstm_cell_const<decltype(BI)::value, 0, true>(ps0 + 4u * 64u * (uint32_t)(tau % kLstmPreRing), c.dyn_slot_cdem, c.dyn_slot_tin, xc, xt, nullptr, h0, c0);
                lstm_cell_const<decltype(BI)::value, 1, false>(0u, -1, 0, 0.f, 0.f, h0, h1, c1);
            }
          }
        };
        switch (bslot) {
            case 0: run(std::integral_constant<int, 0>{}); break;
            case 1: run(std::integral_constant<int, 1>{}); break;
            default: run(std::integral_constant<int, 2>{}); break;
        }
    } else if (pre != nullptr) {
        // every env of the block sits on the same time rows: the exogenous part of W_ih x is shared (helper warp), only the
        // two fed-back inputs are per unit
        const uint32_t ps0 = smem_u32(pre);
#pragma unroll 1
        for (int sidx = 0; sidx < L; ++sidx) {
            const int tau = t - (L - 1) + sidx;
            const float xc = c.dyn_slot_cdem >= 0 ? win_c[(size_t)(tau % ring) * U] : 0.f;
            const float xt = win_t[(size_t)((tau - 1) % ring) * U];
            lstm_cell_pre(ws, ps0 + 4u * 64u * (uint32_t)(tau % kLstmPreRing), c.dyn_slot_cdem, c.dyn_slot_tin, xc, xt, h0, c0);
            lstm_cell<true>(W, ws + 4u * kLstmLayerStride, h0, h1, c1);
        }
    } else
#pragma unroll 1
    for (int sidx = 0; sidx < L; ++sidx) {
        const int tau = t - (L - 1) + sidx;                                 // time step of the non-fed-back inputs
        const float* row = d.table + (size_t)(row0 + tau) * d.Wp + c.dyn_c_inputs;
        float x[kLstmIn];
#pragma unroll
        for (int i = 0; i < kLstmIn; ++i) x[i] = i < c.dyn_n_inputs ? __ldg(row + i) : 0.f;
        const float xc = c.dyn_slot_cdem >= 0 ? win_c[(size_t)(tau % ring) * U] : 0.f;
        const float xt = win_t[(size_t)((tau - 1) % ring) * U];             // indoor temperature is lagged by one step (building.py:3044-3049)
#pragma unroll
        for (int i = 0; i < kLstmIn; ++i) { if (i == c.dyn_slot_cdem) x[i] = xc; if (i == c.dyn_slot_tin) x[i] = xt; }
        if (smem_w) {
            lstm_cell<true>(W, ws, x, h0, c0);
            lstm_cell<true>(W, ws + 4u * kLstmLayerStride, h0, h1, c1);
        } else {
            lstm_cell<false>(W, 0u, x, h0, c0);
            lstm_cell<false>(W + kLstmLayerStride, 0u, h0, h1, c1);
        }
    }
    const float* wl = W + 2 * kLstmLayerStride;
    float y = wl[16];
#pragma unroll
    for (int j = 0; j < kLstmH; ++j) y = fmaf(wl[j], h1[j], y);
#pragma unroll
    for (int j = 0; j < kLstmH; ++j) {
        lst[(size_t)j * U + u] = h0[j]; lst[(size_t)(kLstmH + j) * U + u] = h1[j];
        lst[(size_t)(2 * kLstmH + j) * U + u] = c0[j]; lst[(size_t)(3 * kLstmH + j) * U + u] = c1[j];
    }
    win_t[(size_t)(t % ring) * U] = y;                                       // the prediction replaces the slot (building.py:3027-3028)
    return y * c.tin_range + c.tin_min;                                      // de-normalised (building.py:3031-3037)
}

// ------------------------------------------------------------------------------------------------------------------
// LSTM dynamics on the tensor cores (legacy warp-level path: mma.sync m16n8k8, TF32 operands, FP32 accumulate).
// The scalar cell above is bound by shared-memory operand delivery (one 16-byte weight broadcast per 4 FMAs per lane).  Here a WARP
// whose 32 lanes sit on ONE building treats 16 of its units as the M dimension of D[16 x 8] += A[16 x 8] B[8 x 8]: A = the units'
// hidden vectors, B = an 8 x 8 block of the recurrent (or layer-1 input) matrix.  B fragments come from shared memory in fragment
// order (conflict-free 8-byte loads, 12x fewer shared-memory wavefronts than the scalar cell); each product is done as 3 MMAs on
// hi / lo TF32 splits of both operands (a_lo b_hi + a_hi b_lo + a_hi b_hi: float32-level accuracy, ~2^-21 relative).
// Fragment layout (PTX ISA, m16n8k8 .tf32; g = lane / 4, t = lane % 4):
//   A: a0 (row g, k t)  a1 (row g + 8, k t)  a2 (row g, k t + 4)  a3 (row g + 8, k t + 4)
//   B: b0 (k t, col g)  b1 (k t + 4, col g)            C / D: c0 (g, 2t) c1 (g, 2t + 1) c2 (g + 8, 2t) c3 (g + 8, 2t + 1)
// The K slots are permuted so that slot t of k-tile kt is hidden unit kt * 8 + 2t and slot t + 4 is kt * 8 + 2t + 1: lane (g, t)
// then OWNS hidden units {2t, 2t + 1, 8 + 2t, 9 + 2t} of unit rows g and g + 8 both as producer (the C columns of the i / f / g / o
// tiles of those hidden units) and as consumer (its A registers) - no shuffles between cell steps, h and c stay in registers.
// Fed-back inputs, biases and the shared layer-0 projection `pre` are added in the epilogue in plain float32.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// hi / lo TF32 split of the A registers {ra[j0], rb[j0], ra[j1], rb[j1]}
__device__ __forceinline__ void lstm_a_frag(const float* ra, const float* rb, int j0, int j1, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
    const float v[4] = {ra[j0], rb[j0], ra[j1], rb[j1]};
#pragma unroll
    for (int i = 0; i < 4; ++i) { hi[i] = tf32_rna(v[i]); lo[i] = tf32_rna(v[i] - __uint_as_float(hi[i])); }
}
// acc[q] (gate type q = i, f, g, o; n-tile q * 2 + half) += A (one k-tile, hi / lo) x the fragments of block `blk` (matrix, k-tile)
__device__ __forceinline__ void lstm_mma_ktile(float (&acc)[4][4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t frag /* shared address of the block */,
                                               int half, int lane) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ntile = q * 2 + half;
        uint32_t bh0, bh1, bl0, bl1;
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(bh0), "=r"(bh1) : "r"(frag + 4u * (uint32_t)(((ntile * 2 + 0) * 32 + lane) * 2)));
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(bl0), "=r"(bl1) : "r"(frag + 4u * (uint32_t)(((ntile * 2 + 1) * 32 + lane) * 2)));
        mma_tf32(acc[q], alo, bh0, bh1);
        mma_tf32(acc[q], ahi, bl0, bl1);
        mma_tf32(acc[q], ahi, bh0, bh1);
    }
}
// cell update of the hidden units (half * 8 + 2t, + 1) of rows a, b from the four gate accumulators (+ the epilogue terms already added)
__device__ __forceinline__ void lstm_gate_update(const float (&acc)[4][4], int half, float* ha, float* hb, float* ca, float* cb) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {                       // e: 0 (row a, j0) 1 (row a, j1) 2 (row b, j0) 3 (row b, j1)
        float* hh = (e < 2) ? ha : hb; float* cc = (e < 2) ? ca : cb;
        const int j = half * 2 + (e & 1);               // slot in the lane's 4 owned hidden units: {2t, 2t + 1, 8 + 2t, 9 + 2t}
        const float cn = fmaf(sigmoid_dev(acc[1][e]), cc[j], sigmoid_dev(acc[0][e]) * tanh_dev(acc[2][e]));
        cc[j] = cn;
        hh[j] = sigmoid_dev(acc[3][e]) * tanh_dev(cn);
    }
}

// warp-collective: every lane of the warp sits on building b (units u0 + lane * B), all of them active.  Returns the lane's own
// de-normalised prediction like lstm_update.
template <typename R>
__device__ __forceinline__ float lstm_update_mma(const Dev& d, const UnitCtx<R>& c, const float* Wb /* packed weights, shared */, uint32_t frag /* shared address */,
                                                 const float* pre /* this building's projection ring, shared */, int u, int t, float obs_cool_dem,
                                                 float t_in_dataset, int lane) {
    const int U = d.U, B = d.B;
    const int L = c.dyn_lookback, ring = L + 1;
    float* lst = d.lst;
    float* win_c = lst + (size_t)(4 * kLstmH) * U;                          // [ring][U]
    float* win_t = lst + (size_t)(4 * kLstmH + kLstmMaxLookback + 1) * U;
    // _update_dynamics_input: append the normalised observation of step t (float32 arithmetic), every lane for its own unit
    if (c.dyn_slot_cdem >= 0) win_c[(size_t)(t % ring) * U + u] = (obs_cool_dem - c.cdem_min) / c.cdem_range;
    win_t[(size_t)(t % ring) * U + u] = (t_in_dataset - c.tin_min) / c.tin_range;
    if (t < L) return t_in_dataset;                                         // window not full yet (building.py:2996-2998)
    __syncwarp();                                                           // the lanes read each other's window entries below
    const int g = lane >> 2, tq = lane & 3;
    const int u0 = __shfl_sync(0xffffffffu, u, 0);
    const int jown[4] = {2 * tq, 2 * tq + 1, 8 + 2 * tq, 9 + 2 * tq};
    const int sc = c.dyn_slot_cdem, stn = c.dyn_slot_tin;
    const float* wl = Wb + 2 * kLstmLayerStride;
    float y_mine = 0.f;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int ua = u0 + (pass * 16 + g) * B, ub = ua + 8 * B;           // the units of rows g and g + 8
        float h0a[4], h0b[4], h1a[4], h1b[4], c0a[4], c0b[4], c1a[4], c1b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const size_t j = (size_t)jown[i];
            h0a[i] = lst[j * U + ua]; h0b[i] = lst[j * U + ub];
            h1a[i] = lst[(kLstmH + j) * U + ua]; h1b[i] = lst[(kLstmH + j) * U + ub];
            c0a[i] = lst[(2 * kLstmH + j) * U + ua]; c0b[i] = lst[(2 * kLstmH + j) * U + ub];
            c1a[i] = lst[(3 * kLstmH + j) * U + ua]; c1b[i] = lst[(3 * kLstmH + j) * U + ub];
        }
#pragma unroll 1
        for (int sidx = 0; sidx < L; ++sidx) {
            const int tau = t - (L - 1) + sidx;
            const float xca = sc >= 0 ? win_c[(size_t)(tau % ring) * U + ua] : 0.f, xcb = sc >= 0 ? win_c[(size_t)(tau % ring) * U + ub] : 0.f;
            const float xta = win_t[(size_t)((tau - 1) % ring) * U + ua], xtb = win_t[(size_t)((tau - 1) % ring) * U + ub];   // lagged by one step (building.py:3044-3049)
            const float* pr = pre + 64 * (tau % kLstmPreRing);
            uint32_t ahi[4][4], alo[4][4];
            // ---- layer 0: W_hh0 h0 + pre + the two fed-back input columns ----
            lstm_a_frag(h0a, h0b, 0, 1, ahi[0], alo[0]);
            lstm_a_frag(h0a, h0b, 2, 3, ahi[1], alo[1]);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float acc[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int n0 = q * 16 + half * 8 + 2 * tq;
                    const float p0 = pr[n0], p1 = pr[n0 + 1];
                    const float wc0 = sc >= 0 ? Wb[n0 * 16 + sc] : 0.f, wc1 = sc >= 0 ? Wb[(n0 + 1) * 16 + sc] : 0.f;
                    const float wt0 = Wb[n0 * 16 + stn], wt1 = Wb[(n0 + 1) * 16 + stn];
                    acc[q][0] = fmaf(wt0, xta, fmaf(wc0, xca, p0)); acc[q][1] = fmaf(wt1, xta, fmaf(wc1, xca, p1));
                    acc[q][2] = fmaf(wt0, xtb, fmaf(wc0, xcb, p0)); acc[q][3] = fmaf(wt1, xtb, fmaf(wc1, xcb, p1));
                }
                lstm_mma_ktile(acc, ahi[0], alo[0], frag + 4u * (uint32_t)(0 * kLstmFragBlock), half, lane);
                lstm_mma_ktile(acc, ahi[1], alo[1], frag + 4u * (uint32_t)(1 * kLstmFragBlock), half, lane);
                lstm_gate_update(acc, half, h0a, h0b, c0a, c0b);
            }
            // ---- layer 1: W_ih1 h0(new) + W_hh1 h1 + b1 ----
            lstm_a_frag(h0a, h0b, 0, 1, ahi[0], alo[0]);
            lstm_a_frag(h0a, h0b, 2, 3, ahi[1], alo[1]);
            lstm_a_frag(h1a, h1b, 0, 1, ahi[2], alo[2]);
            lstm_a_frag(h1a, h1b, 2, 3, ahi[3], alo[3]);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float acc[4][4];
                const float* b1 = Wb + kLstmLayerStride + 64 * 32;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int n0 = q * 16 + half * 8 + 2 * tq;
                    acc[q][0] = acc[q][2] = b1[n0]; acc[q][1] = acc[q][3] = b1[n0 + 1];
                }
#pragma unroll
                for (int kt = 0; kt < 4; ++kt) lstm_mma_ktile(acc, ahi[kt], alo[kt], frag + 4u * (uint32_t)((2 + kt) * kLstmFragBlock), half, lane);
                lstm_gate_update(acc, half, h1a, h1b, c1a, c1b);
            }
        }
        // linear head: partial sums over the lane's hidden units, completed over the quad
        float ya = 0.f, yb = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { ya = fmaf(wl[jown[i]], h1a[i], ya); yb = fmaf(wl[jown[i]], h1b[i], yb); }
        ya += __shfl_xor_sync(0xffffffffu, ya, 1); yb += __shfl_xor_sync(0xffffffffu, yb, 1);
        ya += __shfl_xor_sync(0xffffffffu, ya, 2); yb += __shfl_xor_sync(0xffffffffu, yb, 2);
        ya += wl[16]; yb += wl[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const size_t j = (size_t)jown[i];
            lst[j * U + ua] = h0a[i]; lst[j * U + ub] = h0b[i];
            lst[(kLstmH + j) * U + ua] = h1a[i]; lst[(kLstmH + j) * U + ub] = h1b[i];
            lst[(2 * kLstmH + j) * U + ua] = c0a[i]; lst[(2 * kLstmH + j) * U + ub] = c0b[i];
            lst[(3 * kLstmH + j) * U + ua] = c1a[i]; lst[(3 * kLstmH + j) * U + ub] = c1b[i];
        }
        // the owner lane of unit row r (lane = pass * 16 + r) takes its prediction from lane 4 * (r % 8) of this pass
        const int r = lane & 15;
        const float va = __shfl_sync(0xffffffffu, ya, 4 * (r & 7)), vb = __shfl_sync(0xffffffffu, yb, 4 * (r & 7));
        if ((lane >> 4) == pass) y_mine = (r < 8) ? va : vb;
    }
    __syncwarp();
    win_t[(size_t)(t % ring) * U + u] = y_mine;                              // the prediction replaces the slot (building.py:3027-3028)
    return y_mine * c.tin_range + c.tin_min;                                 // de-normalised (building.py:3031-3037)
}

__device__ __forceinline__ void kpi_push(double* a, double x);

// ------------------------------------------------------------------------------------------------------------------
// advance kernel: K consecutive time steps in one launch (cl_step: K = 1, cl_rollout: any K).
//
// Block = P "physics" warps (one thread per unit: building x env, whole envs per block) + ONE helper warp.
//  * physics threads keep parameters, table columns and the unit state in registers for the whole launch; per step they
//    gather their inputs from the TMA-staged time row, take the (prefetched) action, run `unit_step`, publish net / cost /
//    emission to shared memory, pass ONE block barrier, and write rewards (district sums: one thread per (quantity, env), in
//    building order like the reference's sum()).
//  * the helper warp works one step ahead of / behind them: before the barrier of step k it prepares the per-building inputs
//    of step k+1 that do not depend on the env (PV generation: an fp64 division per unit otherwise); after the barrier it
//    streams the observation slab of step k (reference-parity observations are env-independent rows) with
//    16-byte st.global.cs stores WHILE the physics warps already compute step k+1 - the store phase of one step overlaps the
//    arithmetic phase of the next instead of alternating with it.
// The time rows ride a 3-slot TMA ring (cp.async.bulk + mbarrier): the row of step t+3 is requested right after the barrier
// of step t.  `red`, `dsum` and the per-building buffers are double-buffered by step parity, so one barrier per step suffices
// (two when a reward needs the district sum, more for central-agent sums).
// ------------------------------------------------------------------------------------------------------------------
// LSTM districts: block size cap / resident blocks per SM the dynamics instantiation is compiled for (A/B-tested on B200; the LSTM
// loops need ~80 registers, the fp64 thermal physics around them spills either way)
#ifndef CL_DYN_MAXT
#define CL_DYN_MAXT 512
#endif
#ifndef CL_DYN_MINBLOCKS
#define CL_DYN_MINBLOCKS 1
#endif
constexpr int kDynMaxT = CL_DYN_MAXT;
// producer / consumer hand-offs of the DEC instantiation below: mbarriers in shared memory - producers arrive (release), consumers
// poll the phase parity (acquire) WITHOUT arriving, so a consumer never waits for its fellow consumers (a named `bar.sync id, n` is a
// barrier among all n threads: tried first, it was slower than the plain block barrier)
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <typename R, bool THERMAL, bool DYNAMICS, int MAXT, bool WIDE = false, bool KPI = false, bool EVD = false, bool DEC = false>
__global__ void __launch_bounds__(MAXT, (DYNAMICS && MAXT == CL_DYN_MAXT) ? CL_DYN_MINBLOCKS : 1) advance_kernel(Dev d, int t0, int K, const float* __restrict__ actions, float* __restrict__ obs,
                                                        float* __restrict__ reward, float* __restrict__ district, float* __restrict__ trace) {
    extern __shared__ __align__(16) float smf[];
    if (t0 < 0) {
        // device-resident time step (cl_step_device, CUDA-graph replays): a launch past the end of the episode does nothing
        t0 = *reinterpret_cast<const volatile int32_t*>(d.t_dev);
        if (t0 + K > d.T - 1) return;
    }
    const int nt = blockDim.x, tid = threadIdx.x;
    const int np_ = nt - 32;                       // physics threads; the last warp is the helper
    const bool is_helper = tid >= np_;
    const int lane = tid & 31;
    const int B = d.B, epb = d.envs_per_block, Wp = d.Wp;
    // wide districts: CTA `rank` of the cluster owns buildings [b0, b0 + nb) of the cluster's env(s) and columns [k0, k1) of
    // their observation rows; otherwise the block owns whole envs (b0 = 0, nb = B, [k0, k1) = [0, L))
    const int NT = WIDE ? d.tiles : 1;
    const int rank = WIDE ? (int)blockIdx.x % NT : 0;   // == %cluster_ctarank of the 1-D cluster when launched as one
    const int TBs = WIDE ? d.tile_b : B;            // smem stride of per-building arrays
    const int b0 = rank * TBs;
    const int nb = WIDE ? min(TBs, B - b0) : B;
    const int k0 = WIDE ? __ldg(d.tile_k + rank) : 0, k1 = WIDE ? __ldg(d.tile_k + rank + 1) : d.L;
    const int Ltile = k1 - k0;
    const SmemLayout lo = smem_layout(TBs, Wp, WIDE ? d.Lt : d.L, epb, nt, (int)sizeof(R), DYNAMICS ? d.lstm_smem : 0, d.tab_layout, d.fresh_slots, d.ev_n > 0 ? TBs + d.ev_n : d.n_curves, d.kpi_smem);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smf);
    R* scurves = reinterpret_cast<R*>(smf + lo.curves);
    uint8_t* s_clut = reinterpret_cast<uint8_t*>(smf + lo.clut);
    R* s_bsolar = reinterpret_cast<R*>(smf + lo.bsolar);          // [2][B] PV generation of the step, per building
    float* s_rows = smf + lo.rows;
    int32_t* s_tcol = reinterpret_cast<int32_t*>(smf + lo.tcol);
    float* s_dsum = smf + lo.dsum;
    float* s_rsum = smf + lo.rsum;
    float* s_dynbuf = smf + lo.dynbuf;
    float* s_wpart = smf + lo.wpart;
    double* s_rpart = reinterpret_cast<double*>(smf + lo.rpart);
    constexpr bool kpi = KPI && !WIDE && !DYNAMICS;     // a separate instantiation: the plain step kernel carries none of this code
    double* s_kacc = reinterpret_cast<double*>(smf + lo.kpi_acc);      // [CL_NKPI_UNIT][nt]
    double* s_knws = reinterpret_cast<double*>(smf + lo.kpi_nws);      // [2][nt]
    double* s_kenv = reinterpret_cast<double*>(smf + lo.kpi_env);      // [epb][2][CL_NKPI_ENV]
    const int e0 = (WIDE ? (int)blockIdx.x / NT : (int)blockIdx.x) * epb;
    const int n_env = min(epb, d.E - e0);
    const int n_units = n_env * nb;
    const bool active = tid < n_units;
    // thread -> unit: env-major (building fastest: coalesced state / action / reward slices) except for LSTM districts, where
    // building-major keeps the lanes of a warp on ONE building so that its LSTM weights are broadcast loads
    const int bl = DYNAMICS ? tid / n_env : tid % nb;       // building inside the tile
    const int e_l = DYNAMICS ? tid - bl * n_env : tid / nb;
    const int b = b0 + bl;
    const int ul = e_l * nb + bl;                  // unit slot inside the block's shared-memory arrays (always env-major)
    const int e = e0 + e_l, u = e * B + b;
    const bool uniform = d.uniform_start != 0;
    // observations with action-dependent (DYN) columns and an observation table: per-env row images in shared memory
    const bool fresh_tab = obs != nullptr && uniform && !d.stale && d.obs_tab != nullptr && d.fresh_slots;
    const bool want_dyn = (!d.stale && obs != nullptr) && !fresh_tab;      // general writer (per-env windows, no table)
    const bool tmpl_path = obs != nullptr && uniform && d.stale && !d.obs_state;      // (per-env state columns: general writer)
    const bool need_dsum = (d.reward_id == CL_REWARD_MARL || d.reward_id == CL_REWARD_ELECTRIC_VEHICLES) && reward != nullptr;
    constexpr bool has_ev = EVD && !WIDE && !DYNAMICS;                  // chargers / washing machines: a separate instantiation (like KPI)
    // DEC: no block-wide barrier inside the step.  When nothing in the step reads the district sum (default-type rewards, decentralised
    // agents, reference-parity observations, one episode window) the physics warps do not depend on each other: a step barrier only
    // makes every warp wait for the slowest one of THAT step (divergent charge / discharge paths: 15 % of the stall samples).  Here the
    // barrier is split in time: a warp ARRIVES when its step-k results are in shared memory, goes on with the physics of step k + 1 and
    // only then WAITS for the arrivals of step k - by then a whole step old - to add up its share of the district sums of step k.
    // mbarriers in shared memory (s_bar + 5 ..), two of each by step parity p = k & 1:
    //   A[p] (5, 6)   helper -> physics   per-building inputs of step k are in shared memory                    1 arrival per phase
    //   R[p] (7, 8)   physics -> all      step k: red[p] written, row t and the inputs of parity p read         1 arrival per physics warp
    //   S[p] (9, 10)  physics -> physics  the sums of step k have been read out of red[p] (written again at step k + 2)   "
    constexpr bool dec = DEC && !WIDE && !DYNAMICS && !KPI && !EVD;
    const bool fused_reward = reward != nullptr && d.reward_id >= 0;
    const bool central_sync = fused_reward && d.central;
    const int Rdim = d.central ? 1 : B;
    const uint32_t row_bytes = (uint32_t)Wp * sizeof(float);

    const bool tab_path = (tmpl_path && d.obs_tab != nullptr) || fresh_tab;   // the host only sets obs_tab when every tile range is 16-byte aligned
    const int n_slots = fresh_tab ? n_env : 1;                     // row images per buffer
    const uint32_t slab_bytes = (uint32_t)Ltile * sizeof(float);
    const bool coupled = WIDE && d.coupled;                        // cross-tile sums needed inside the step
    if (uniform && tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(s_bar + i, 1);
        if (dec) { for (int i = 5; i < 11; ++i) mbar_init(s_bar + i, i >= 7 ? (np_ >> 5) : 1); }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int i = 0; i < 3 && i <= K; ++i) {     // rows t0 .. t0+2 (row t0+K is the last one any step needs)
            mbar_expect_tx(s_bar + i, row_bytes);
            tma_load_1d(s_rows + i * Wp, d.table + (size_t)(d.start0 + t0 + i) * Wp, row_bytes, s_bar + i);
        }
        if (tab_path) {                             // observation row of step 0 (table row start0 + t0 + 1) into buffer 0
            mbar_expect_tx(s_bar + 3, (uint32_t)n_slots * slab_bytes);
            for (int le = 0; le < n_slots; ++le)
                tma_load_1d(smf + lo.tmpl + (size_t)le * lo.Lp, d.obs_tab + (size_t)(d.start0 + t0 + 1) * d.obs_pitch + k0, slab_bytes, s_bar + 3);
        }
    }
    // block-wide staging: template columns and the battery curves of every building (dynamic indexing -> shared memory)
    if (tmpl_path && !tab_path) for (int k = tid; k < Ltile; k += nt) s_tcol[k] = __ldg(d.tcol + k0 + k);
    {
        // per building: [PE_X 8][PE_Y 8][CP_X 8][CP_Y 8][PE_RW 8][CP_RW 8] (SmemCurves): x entries beyond the curve's points are
        // +inf, RW[k] = refined reciprocal of the segment width x[k+1] - x[k]
        const auto* P = PSel<R>::p(d) + CL_P_PE_X0 * B;
        // distinct tables of the district, or one per building of the tile - followed by the vehicles' battery curves (tables nb ..)
        const int ncv = (d.n_curves > 0 ? d.n_curves : nb) + d.ev_n;
        const int nbt = ncv - d.ev_n;
        auto cpar = [&](int ci, int jj) -> R {          // entry jj of [PE_X 8][PE_Y 8][CP_X 8][CP_Y 8] of table ci
            if (ci >= nbt) return (R)__ldg(d.ev_pd + (size_t)(ci - nbt) * CL_NPARAM + CL_P_PE_X0 + jj);
            return (R)__ldg(P + jj * B + (d.n_curves > 0 ? __ldg(d.curve_rep + ci) : b0 + ci));
        };
        auto cnum = [&](int ci, int which) -> int {
            if (ci >= nbt) return __ldg(d.ev_ip + (ci - nbt) * 2 + which);
            return __ldg(d.ip + (which ? CL_IP_CP_N : CL_IP_PE_N) * B + (d.n_curves > 0 ? __ldg(d.curve_rep + ci) : b0 + ci));
        };
        for (int i = tid; i < ncv * kCurveTab; i += nt) {
            const int ci = i / kCurveTab, j = i % kCurveTab;
            R v;
            if (j < 4 * CL_MAX_CURVE) {
                v = cpar(ci, j);
                const int which = j >> 4, k = j & 7;
                if (!(j & 8) && k >= cnum(ci, which)) v = Num<R>::inf();
            } else {
                const int which = (j - 4 * CL_MAX_CURVE) >> 3, k = j & 7;
                const int n = cnum(ci, which);
                v = (R)0;
                if (k + 1 < n) v = make_divisor(cpar(ci, which * 16 + k + 1) - cpar(ci, which * 16 + k)).r;
            }
            scurves[i] = v;
        }
        if (d.curve_lut != nullptr) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(d.curve_lut);       // [B + ev_n] indices: buildings, then vehicles
            for (int i = tid; i < ncv * kCurveLutFloats; i += nt) {
                const int ci = i / kCurveLutFloats, j = i - ci * kCurveLutFloats;
                const int bb = ci >= nbt ? B + (ci - nbt) : (d.n_curves > 0 ? __ldg(d.curve_rep + ci) : b0 + ci);
                reinterpret_cast<uint32_t*>(s_clut)[i] = __ldg(src + (size_t)bb * kCurveLutFloats + j);
            }
        }
    }
    UnitCtx<R> c;
    UnitState<R> s;
    int start_e = 0;
    int dyn_lo = 0, dyn_hi = 0;                    // this building's (column, cl_dyn slot) pairs in d.dyn_cols
    RawActions act_next = {};
    if (active) {
        if (fresh_tab) { dyn_lo = __ldg(d.dyn_off + b); dyn_hi = __ldg(d.dyn_off + b + 1); }
        load_ctx<R, THERMAL>(d, b, c);
        load_state<R, THERMAL>(d, u, s);
        start_e = __ldg(d.start + e);
        fetch_actions<R, THERMAL>(d, c, actions + (size_t)e * d.A, act_next);
    }
    const int my_curve = !active ? 0 : (d.n_curves > 0 ? __ldg(d.curve_id + b) : bl);
    const SmemCurves<R> curves = {scurves + my_curve * kCurveTab, d.curve_nmax,
                                  d.curve_lut != nullptr ? s_clut + my_curve * (2 * kCurveLutStride) : nullptr};
    const float* lstm_w = nullptr;
    if (DYNAMICS) {
        if (d.lstm_smem) {                            // stage every building's packed LSTM weights in shared memory (16-byte copies)
            float4* dst = reinterpret_cast<float4*>(smf + lo.lstm);
            const float4* src = reinterpret_cast<const float4*>(d.lstm_w);
            for (int i = tid; i < B * (kLstmStride / 4); i += nt) dst[i] = __ldg(src + i);
            lstm_w = smf + lo.lstm + (size_t)(active ? b : 0) * kLstmStride;
        } else {
            lstm_w = d.lstm_w + (size_t)(active ? b : 0) * kLstmStride;
        }
    }
    if (kpi) {
#pragma unroll
        for (int j = 0; j < CL_NKPI_UNIT; ++j) s_kacc[j * nt + tid] = 0.0;
        for (int i = tid; i < n_env * 2 * CL_NKPI_ENV; i += nt) s_kenv[i] = d.kpi_env[(size_t)e0 * 2 * CL_NKPI_ENV + i];
    }
    __syncthreads();   // mbarriers initialised (visible to every waiter), curves / tcol / LSTM weights staged
    if (DYNAMICS && d.lstm_smem == 2) {
        // tensor-core operand fragments of W_hh0, W_ih1, W_hh1 (lstm_update_mma), from the packed weights staged above:
        // block (matrix m, k-tile kt), n-tile, {hi, lo}, lane (g, t): {W[n-tile * 8 + g][kt * 8 + 2t], W[..][kt * 8 + 2t + 1]}
        float* fr = smf + lo.lstm_frag;
        for (int i = tid; i < B * 6 * 8 * 32; i += nt) {
            const int ln = i & 31, ntile = (i >> 5) & 7, blk = (i >> 8) % 6, bb = i / (6 * 8 * 32);
            const int m = blk >> 1, kt = blk & 1, gg = ln >> 2, tt = ln & 3;
            const float* Wm = smf + lo.lstm + (size_t)bb * kLstmStride + (m == 0 ? 64 * 16 : kLstmLayerStride + (m == 2 ? 64 * 16 : 0));
            const float w0 = Wm[(ntile * 8 + gg) * 16 + kt * 8 + 2 * tt], w1 = Wm[(ntile * 8 + gg) * 16 + kt * 8 + 2 * tt + 1];
            const uint32_t h0 = tf32_rna(w0), h1 = tf32_rna(w1);
            float* dst = fr + (size_t)bb * kLstmFragFloats + (size_t)blk * kLstmFragBlock;
            dst[((ntile * 2 + 0) * 32 + ln) * 2] = __uint_as_float(h0); dst[((ntile * 2 + 0) * 32 + ln) * 2 + 1] = __uint_as_float(h1);
            dst[((ntile * 2 + 1) * 32 + ln) * 2] = __uint_as_float(tf32_rna(w0 - __uint_as_float(h0)));
            dst[((ntile * 2 + 1) * 32 + ln) * 2 + 1] = __uint_as_float(tf32_rna(w1 - __uint_as_float(h1)));
        }
        __syncthreads();
    }

    // LSTM layer-0 input projections shared by all envs of the block (uniform episode windows, weights in shared memory):
    // pre[bb][tau % ring][r] = bias0[r] + W_ih0[r][:] . x(row of time step tau); fed-back slots hold 0 in the table
    const bool use_pre = DYNAMICS && uniform && d.lstm_smem;
    // tensor-core cell: needs the fragments, the shared projections, and whole warps on one building (building-major mapping with a
    // multiple of 32 envs in this block)
    const bool lstm_mma = use_pre && d.lstm_smem == 2 && (n_env & 31) == 0;
    float* s_pre = smf + lo.lstm_pre;
    auto project_row = [&](int bb, int tau, const float* rowp, int r) {
        const float* Wb = smf + lo.lstm + (size_t)bb * kLstmStride;
        const int nin = __ldg(d.ip + CL_IP_DYN_N_INPUTS * B + bb);
        const float* xin = rowp + __ldg(d.ip + CL_IP_DYN_C_INPUTS * B + bb);
        float acc = Wb[64 * 32 + r];
        for (int i = 0; i < nin; ++i) acc = fmaf(Wb[r * 16 + i], xin[i], acc);
        s_pre[((size_t)bb * kLstmPreRing + (tau % kLstmPreRing)) * 64 + r] = acc;
    };
    if (use_pre) {
        // rows t0 - 11 .. t0 straight from the table (later rows are projected by the helper warp one step ahead)
        for (int idx = tid; idx < B * kLstmMaxLookback * 64; idx += nt) {
            const int r = idx & 63, rest = idx >> 6;
            const int bb = rest / kLstmMaxLookback, tau = t0 - (rest - bb * kLstmMaxLookback);
            if (tau >= 0 && (__ldg(d.ip + CL_IP_FLAGS * B + bb) & CL_F_DYNAMICS))
                project_row(bb, tau, d.table + (size_t)(d.start0 + tau) * Wp, r);
        }
    }

    // per-building PV generation of time row `rowp` -> dst[b]  (building.py:2554; the same value for every env of the block)
    const Divisor<R> div1000 = make_divisor((R)1000);
    auto building_inputs = [&](const float* rowp, R* dst) {
        const auto* P = PSel<R>::p(d);
        for (int bb = lane; bb < nb; bb += 32) {
            const R pv = (R)__ldg(P + CL_P_PV_NOMINAL_POWER * B + b0 + bb);
            dst[bb] = -dvr(pv * (R)rowp[__ldg(d.ip + CL_IP_C_SOLAR * B + b0 + bb)], div1000);
        }
    };
    if (uniform) {
        if (is_helper) {
            mbar_wait(s_bar + 0, 0u);
            building_inputs(s_rows, s_bsolar);
        }
        __syncthreads();
    }

#ifdef CL_PHASE_TIMING
#define CL_STAMP(i) do { if (blockIdx.x == 1 && tid == 32 && trace) trace[k * 8 + (i)] = (float)(clock64() - clk0); } while (0)
    const long long clk0 = clock64();
#else
#define CL_STAMP(i)
#endif
    for (int k = 0; k < K; ++k) {
        const int t = t0 + k;
        const int pb = k & 1;
        CL_STAMP(0);
        const int slot_t = k % 3, slot_n = (k + 1) % 3;
        RowRef row;
        const float* row_next = nullptr;
        if (uniform) {
            mbar_wait(s_bar + slot_t, (uint32_t)((k / 3) & 1));
            mbar_wait(s_bar + slot_n, (uint32_t)(((k + 1) / 3) & 1));
            row.g = s_rows + slot_t * Wp; row.s = smem_u32(row.g);
            row_next = s_rows + slot_n * Wp;
        } else {
            row.g = d.table + (size_t)(start_e + t) * Wp; row.s = 0u;
        }
        CL_STAMP(1);
        float* red = smf + lo.red + pb * 3 * nt;

        if (is_helper) {
            // ---------------- helper warp ----------------
            if (uniform && k + 1 < K) building_inputs(row_next, s_bsolar + ((k + 1) & 1) * TBs);
            if (use_pre && k + 1 < K) {
                for (int bb = 0; bb < B; ++bb) {
                    if (!(__ldg(d.ip + CL_IP_FLAGS * B + bb) & CL_F_DYNAMICS)) continue;
                    project_row(bb, t + 1, row_next, lane); project_row(bb, t + 1, row_next, lane + 32);
                }
            }
            if (dec) {
                __syncwarp();
                if (lane == 0) {
                    if (k == 0) mbar_arrive(s_bar + 5);                         // A[0]: the prologue's inputs of step 0
                    if (k + 1 < K) mbar_arrive(s_bar + 5 + ((k + 1) & 1));      // A: the inputs of step k + 1 are in place
                }
            } else {
            if (coupled) cluster_sync_all(); else __syncthreads();             // S1
            if (need_dsum) __syncthreads();                                    // S2
            if (central_sync) { if (WIDE) cluster_sync_all(); else { if (k > 0) __syncthreads(); __syncthreads(); } }
            }
            if (tab_path) {
                // observation slab of step k straight from the precomputed table: buffer pb holds columns [k0, k1) of table row
                // start0 + t + 1 (requested one step ago) - ONE image shared by every env (reference-parity rows are env-independent) or,
                // with action-dependent columns, one image per env that the physics threads have patched before S1.  Patch the outage
                // columns, bulk-store the image(s) into the env rows, then request the row of step k + 1 into the other buffer once
                // its previous stores have read it
                float* ok = obs + (size_t)k * d.E * d.L + (size_t)e0 * d.L + k0;
                float* tmpl = smf + lo.tmpl + (size_t)pb * n_slots * lo.Lp;
                mbar_wait(s_bar + 3 + pb, (uint32_t)((k >> 1) & 1));
                if (d.has_outage && d.n_out_cols > 0) {
                    for (int i = lane; i < d.n_out_cols; i += 32) {
                        const int kk = __ldg(d.out_cols + i);
                        if (kk >= k0 && kk < k1) {
                            const int bb = __ldg(d.desc + kk).w;
                            float v = __ldg(d.outage + bb * d.T + t + 1);
                            if (d.obs_t) v = transform_obs(d.obs_t, kk, v);
                            for (int le = 0; le < n_slots; ++le) tmpl[(size_t)le * lo.Lp + kk - k0] = v;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                }
                if (lane == 0) {
                    for (int le = 0; le < n_env; ++le) tma_store_1d(ok + (size_t)le * d.L, tmpl + (fresh_tab ? (size_t)le * lo.Lp : 0), slab_bytes);
                    tma_store_commit();
                    if (k + 1 < K) {
                        tma_store_wait_read<1>();                               // the stores of step k - 1 have read the other buffer
                        float* other = smf + lo.tmpl + (size_t)(pb ^ 1) * n_slots * lo.Lp;
                        mbar_expect_tx(s_bar + 3 + (pb ^ 1), (uint32_t)n_slots * slab_bytes);
                        for (int le = 0; le < n_slots; ++le)
                            tma_load_1d(other + (size_t)le * lo.Lp, d.obs_tab + (size_t)(d.start0 + t + 2) * d.obs_pitch + k0, slab_bytes, s_bar + 3 + (pb ^ 1));
                    }
                }
                __syncwarp();
            } else if (tmpl_path) {
                // observation slab of step k: every env row of the block equals the row gathered from time row t+1
                float* ok = obs + (size_t)k * d.E * d.L + (size_t)e0 * d.L + k0;
                const int L = d.L;
                auto gather = [&](int j) -> float {
                    const int cc = s_tcol[j];
                    float v;
                    if (cc >= 0) v = row_next[cc];
                    else if (cc == -1) v = 0.f;
                    else v = d.has_outage ? __ldg(d.outage + (-2 - cc) * d.T + t + 1) : 0.f;
                    return d.obs_t ? transform_obs(d.obs_t, k0 + j, v) : v;
                };
                if (((L | k0 | Ltile) & 3) == 0) {
                    // build the row once in shared memory, then one TMA bulk store (cp.async.bulk shared -> global) per env row:
                    // the copy engine moves the slab, the SM's load/store path stays free for the physics warps
                    float* tmpl = smf + lo.tmpl + pb * lo.Lp;
                    tma_store_wait_read<1>();                                   // the stores issued two steps ago have read this buffer
                    for (int j = lane; j < Ltile; j += 32) tmpl[j] = gather(j);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the async proxy
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t bytes = (uint32_t)Ltile * sizeof(float);
                        for (int le = 0; le < n_env; ++le) tma_store_1d(ok + (size_t)le * L, tmpl, bytes);
                        tma_store_commit();
                    }
                } else {
                    for (int j = lane; j < Ltile; j += 32) {
                        const float v = gather(j);
                        for (int le = 0; le < n_env; ++le) __stcs(ok + (size_t)le * L + j, v);
                    }
                }
            } else if (obs != nullptr && want_dyn && k + 1 < K) {
                __syncthreads();
            }
            if (has_ev && k + 1 < K) __syncthreads();
            if (dec) {
                mbar_wait(s_bar + 7 + pb, (uint32_t)((k >> 1) & 1));            // R: every physics warp has finished step k
                if (lane == 0 && k + 3 <= K) {
                    // nobody reads row t any more: its slot takes the row of step t + 3
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(s_bar + slot_t, row_bytes);
                    tma_load_1d(s_rows + slot_t * Wp, d.table + (size_t)(d.start0 + t + 3) * Wp, row_bytes, s_bar + slot_t);
                }
            }
            continue;
        }

        // ---------------- physics warps ----------------
        UnitResult<R> o;
        RewardIn ri;
        ChargerInfo chi[CL_MAX_CHARGERS_PER_BUILDING];
        int n_chi = 0;
        double cc_penalty = 0.0;          // charging-constraint violation x coefficient of this step (ev_reward)
        if (dec) {
            mbar_wait(s_bar + 5 + pb, (uint32_t)((k >> 1) & 1));                       // A: this step's per-building inputs (helper warp, one step ahead)
        }
        if (active) {
            UnitInputs<R> in;
            load_inputs<R, THERMAL>(d, c, row, b, t, in, uniform);
            if (uniform) in.solar = s_bsolar[pb * TBs + bl];
            apply_actions<R, THERMAL>(d, c, act_next, in);
            if (k + 1 < K) fetch_actions<R, THERMAL>(d, c, actions + ((size_t)(k + 1) * d.E + e) * d.A, act_next);   // prefetch
            if (DYNAMICS && (c.p.flags & CL_F_DYNAMICS) && t > c.dyn_lookback) {
                // partial-load control is live once the input window is full (building.py:3108, 3144)
                in.control_cooling_demand = (c.a_cd >= 0 || c.a_coh >= 0);
                in.control_heating_demand = (c.a_hd >= 0 || c.a_coh >= 0);
            }
            if (has_ev) {
                // chargers, then washing machines (appended to the priority list, building.py:1582-1604; they neither use nor change
                // the building's flexibility, so running them before the building's own devices gives the same values)
                float ch_tot = 0.f, wm_tot = 0.f;
                const float* arow = actions + ((size_t)k * d.E + e) * d.A;
                const size_t EV = (size_t)d.E * d.ev_n;
                const int tab0 = d.n_curves > 0 ? d.n_curves : nb;     // vehicles' curve tables follow the buildings'
                const int ch0 = __ldg(d.ch_off + b), ch1 = __ldg(d.ch_off + b + 1);
                double act_c[CL_MAX_CHARGERS_PER_BUILDING];
#pragma unroll
                for (int j = 0; j < CL_MAX_CHARGERS_PER_BUILDING; ++j) {
                    act_c[j] = 0.0;
                    const int slot = ch0 + j < ch1 ? __ldg(d.ch_action + ch0 + j) : -1;
                    if (slot >= 0) {
                        float av = __ldg(arow + slot);
                        if (d.act_range != nullptr) av = av * __ldg(d.act_range + slot) + __ldg(d.act_low + slot);
                        act_c[j] = (double)av;
                    }
                }
                const int cci = d.cc_n > 0 ? __ldg(d.cc_index + b) : -1;
                if (cci >= 0) {
                    // Building._apply_charging_constraints_to_actions (building.py:894-982): python-float arithmetic, `sum()` = 0 + x0 + x1 ..
                    const double* lim = d.cc_limits + (size_t)cci * (1 + CL_MAX_PHASES);
                    const int32_t* mem = d.cc_members + (size_t)cci * CL_MAX_PHASES * CL_MAX_CHARGERS_PER_BUILDING;
                    float* stt = d.cc_state + (size_t)e * (d.cc_n * CL_CC_SLOTS) + (size_t)cci * CL_CC_SLOTS;
                    double req[CL_MAX_CHARGERS_PER_BUILDING], scl[CL_MAX_CHARGERS_PER_BUILDING], maxp[CL_MAX_CHARGERS_PER_BUILDING];
                    bool pos[CL_MAX_CHARGERS_PER_BUILDING];
                    bool any = false;
                    double total = 0.0;
#pragma unroll
                    for (int j = 0; j < CL_MAX_CHARGERS_PER_BUILDING; ++j) {
                        const bool has = ch0 + j < ch1 && __ldg(d.ch_action + ch0 + j) >= 0;
                        maxp[j] = has ? __ldg(d.ch_pd + (size_t)(ch0 + j) * CL_NCHP + CL_CH_MAX_C) : 0.0;
                        pos[j] = has && act_c[j] > 0.0 && maxp[j] > 0.0;
                        req[j] = pos[j] ? act_c[j] * maxp[j] : 0.0;
                        scl[j] = 1.0;
                        if (pos[j]) { total = total + req[j]; any = true; }
                    }
                    const double blim = __ldg(lim);
                    double viol = 0.0, head_b = blim, head_p[CL_MAX_PHASES];
#pragma unroll
                    for (int ph = 0; ph < CL_MAX_PHASES; ++ph) head_p[ph] = __ldg(lim + 1 + ph);
                    if (any) {
                        if (blim == blim && blim >= 0.0 && total > blim) {
                            const double sc = blim == 0.0 ? 0.0 : blim / total;
#pragma unroll
                            for (int j = 0; j < CL_MAX_CHARGERS_PER_BUILDING; ++j) if (pos[j]) scl[j] *= sc;
                            viol += total - blim;
                        }
                        for (int ph = 0; ph < CL_MAX_PHASES; ++ph) {
                            const double pl = __ldg(lim + 1 + ph);
                            if (!(pl == pl) || pl < 0.0) continue;
                            double psum = 0.0;
                            for (int i = 0; i < CL_MAX_CHARGERS_PER_BUILDING; ++i) {
                                const int m = __ldg(mem + ph * CL_MAX_CHARGERS_PER_BUILDING + i);
                                if (m < 0) break;
                                if (pos[m]) psum = psum + req[m] * scl[m];
                            }
                            if (psum > pl) {
                                const double ps = pl == 0.0 ? 0.0 : pl / psum;
                                for (int i = 0; i < CL_MAX_CHARGERS_PER_BUILDING; ++i) {
                                    const int m = __ldg(mem + ph * CL_MAX_CHARGERS_PER_BUILDING + i);
                                    if (m < 0) break;
                                    if (pos[m]) scl[m] *= ps;
                                }
                                viol += psum - pl;
                            }
                        }
                        double used = 0.0;
#pragma unroll
                        for (int j = 0; j < CL_MAX_CHARGERS_PER_BUILDING; ++j) {
                            const bool has = ch0 + j < ch1 && __ldg(d.ch_action + ch0 + j) >= 0;
                            if (!has || !(act_c[j] > 0.0)) continue;
                            if (maxp[j] <= 0.0) { act_c[j] = 0.0; continue; }
                            const double scaled = req[j] * scl[j];
                            used = used + scaled;
                            act_c[j] = fmax(0.0, fmin(act_c[j], scaled / maxp[j]));
                            req[j] = scaled;                                  // (kept for the phase headrooms below)
                        }
                        head_b = blim - used;
                        for (int ph = 0; ph < CL_MAX_PHASES; ++ph) {
                            double up = 0.0;
                            for (int i = 0; i < CL_MAX_CHARGERS_PER_BUILDING; ++i) {
                                const int m = __ldg(mem + ph * CL_MAX_CHARGERS_PER_BUILDING + i);
                                if (m < 0) break;
                                if (pos[m]) up = up + req[m];
                            }
                            head_p[ph] = head_p[ph] - up;
                        }
                    }
                    const double pen = viol * (double)c.p.hours;
                    stt[0] = (float)head_b;
#pragma unroll
                    for (int ph = 0; ph < CL_MAX_PHASES; ++ph) stt[1 + ph] = (float)head_p[ph];
                    stt[CL_CC_SLOTS - 1] = (float)pen;
                    if ((__ldg(d.cc_flags + cci) & 1) && pen > 0.0) cc_penalty = pen * (double)d.rp[0];
                }
                for (int kk = ch0; kk < ch1; ++kk) {
                    const int4 cc = __ldg(reinterpret_cast<const int4*>(d.ch_cols) + kk);
                    const bool conn = row[cc.x] > 0.f;
                    const int v = conn ? (int)row[cc.y] : 0;
                    const int slot = __ldg(d.ch_action + kk);
                    const double a = act_c[kk - ch0];
                    const double* qp = d.ch_pd + (size_t)kk * CL_NCHP;
                    ChargerParams<R> q;
                    q.max_c = (R)__ldg(qp + CL_CH_MAX_C); q.min_c = (R)__ldg(qp + CL_CH_MIN_C); q.max_d = (R)__ldg(qp + CL_CH_MAX_D);
                    q.min_d = (R)__ldg(qp + CL_CH_MIN_D); q.eff = (R)__ldg(qp + CL_CH_EFF);
                    q.c_n = (int)__ldg(qp + CL_CH_C_N); q.d_n = (int)__ldg(qp + CL_CH_D_N); q.curves = qp + CL_CH_C_X0;
                    const double* ep = d.ev_pd + (size_t)v * CL_NPARAM;
                    const size_t i = (size_t)e * d.ev_n + v;
                    BuildingParams<R> evp;
                    UnitState<R> evs;
                    evp.bat_capacity = (R)__ldg(ep + CL_P_BAT_CAPACITY); evp.bat_pnom = (R)__ldg(ep + CL_P_BAT_NOMINAL_POWER);
                    evp.bat_loss = (R)__ldg(ep + CL_P_BAT_LOSS); evp.bat_clc = (R)__ldg(ep + CL_P_BAT_CLC); evp.bat_dod = (R)__ldg(ep + CL_P_BAT_DOD);
                    evp.ratio = (R)__ldg(ep + CL_P_TIME_STEP_RATIO); evp.hours = (R)__ldg(ep + CL_P_HOURS_PER_STEP); evp.flags = 0;
                    evp.pe_n = __ldg(d.ev_ip + v * 2); evp.cp_n = __ldg(d.ev_ip + v * 2 + 1);
                    derive_params(evp);
                    // Battery.charge starts from the soc[t-1] entry - the soc[0] entry at t == 0 (energy_model.py:662-666, 1046)
                    evs.soc_b = (R)(t == 0 ? d.ev_sf[EV + i] : d.ev_sf[i]); evs.cap_deg = (R)d.ev_sd[i]; evs.rte_b = (R)d.ev_sd[EV + i];
                    evs.soc_cs = evs.soc_hs = evs.soc_ds = (R)0;
                    const SmemCurves<R> evc = {scurves + (size_t)(tab0 + v) * kCurveTab, d.curve_nmax,
                                               d.curve_lut != nullptr ? s_clut + (size_t)(tab0 + v) * (2 * kCurveLutStride) : nullptr};
                    float kwh; bool charged;
                    const float ecc = charger_step<R>(q, a, conn, evp, evc, d.ev_flag[i] == 0, evs, (double)evp.hours, kwh, charged);
                    if (charged) {
                        d.ev_sf[EV + i] = (float)evs.soc_b; d.ev_sd[i] = (double)evs.cap_deg; d.ev_sd[EV + i] = (double)evs.rte_b; d.ev_flag[i] = 1;
                    }
                    ch_tot = ch_tot + ecc;                                          // `0 + float32 + ...` in charger order (building.py:2654-2661)
                    if (n_chi < CL_MAX_CHARGERS_PER_BUILDING) {
                        ChargerInfo& ci = chi[n_chi++];
                        ci.conn = conn; ci.kwh = conn ? kwh : 0.f; ci.soc_now = d.ev_sf[EV + i];
                        ci.soc_prev = t == 0 ? __ldg(ep + CL_P_BAT_INITIAL_SOC) : (double)d.ev_sf[i];
                        ci.cap = __ldg(ep + CL_P_BAT_CAPACITY); ci.min_cap = (1.0 - __ldg(ep + CL_P_BAT_DOD)) * ci.cap;
                        ci.req = (double)row[cc.z]; ci.hrs = (double)row[cc.w]; ci.max_c = (double)q.max_c; ci.max_d = (double)q.max_d;
                    }
                }
                for (int kk = __ldg(d.wm_off + b); kk < __ldg(d.wm_off + b + 1); ++kk) {
                    // WashingMachine.next_time_step / start_cycle (energy_model.py:1289-1327)
                    const int4 wc = __ldg(reinterpret_cast<const int4*>(d.wm_cols) + kk);
                    const int st = (int)row[wc.x], en = (int)row[wc.y];
                    const size_t wi = (size_t)e * d.wm_n + kk;
                    bool init = d.wm_flag[wi] != 0;
                    if (t > 0) {
                        const float* prow = d.table + (size_t)(start_e + t - 1) * Wp;   // a new window re-arms the machine
                        if ((int)__ldg(prow + wc.x) != st || (int)__ldg(prow + wc.y) != en) init = false;
                    }
                    const int slot = __ldg(d.wm_action + kk);
                    if (slot >= 0) {
                        float av = __ldg(arow + slot);
                        if (d.act_range != nullptr) av = av * __ldg(d.act_range + slot) + __ldg(d.act_low + slot);
                        if (!init && av > 0.f && st != -1 && en != -1 && st <= t && t <= en && (int)row[wc.w] > 0) {
                            // every entry of the load profile lands in ec[t]; entries whose step t + offset is past the episode are skipped
                            const int len = (int)row[wc.w], left = d.T - t;
                            wm_tot = wm_tot + row[len > left ? wc.z + left : wc.z];
                            init = true;
                        }
                    }
                    d.wm_flag[wi] = init ? 1 : 0;
                }
                in.chargers_ec = (R)ch_tot; in.machines_ec = (R)wm_tot;
            }
            CL_STAMP(2);
            unit_step<R, THERMAL, has_ev>(c.p, curves, t, in, s, o);
            CL_STAMP(3);
            float t_in = row[c.c_tin];
            if (DYNAMICS && (c.p.flags & CL_F_DYNAMICS)) {
                const float cd = (float)(o.e_from_cool + fabs(rmin(o.eb_cs, (R)0)));
                // (the constant-operand cells are compiled into the regular dynamics instantiation only: build time)
#ifdef CL_LSTM_CONST            // experiment (profiles/README.md): constant-bank weight operands were 4.5x SLOWER than shared memory
                constexpr bool kConstW = DYNAMICS && MAXT == kDynMaxT;
#else
                constexpr bool kConstW = false;
#endif
                if (lstm_mma) {
                    // (whole warps on one building, all lanes active: see lstm_mma above)
                    t_in = lstm_update_mma<R>(d, c, lstm_w, smem_u32(smf + lo.lstm_frag + (size_t)b * kLstmFragFloats), s_pre + (size_t)b * kLstmPreRing * 64,
                                              u, t, cd, t_in, lane);
                } else
                t_in = lstm_update<R, kConstW>(d, c, lstm_w, u, t, start_e, cd, t_in, use_pre ? s_pre + (size_t)b * kLstmPreRing * 64 : nullptr, b);
            }
            if (dec && k >= 2) mbar_wait(s_bar + 9 + pb, (uint32_t)(((k >> 1) - 1) & 1));   // S: the sums of step k - 2 have left red[pb]
            red[ul] = (float)o.net;
            red[nt + ul] = (float)o.cost;
            red[2 * nt + ul] = (float)o.emission;
            if (fused_reward) reward_inputs<R, THERMAL>(d, c, s, o, row, t_in, ri);     // everything the reward needs from row t
            if (kpi) {
                // CityLearnEnv.evaluate()'s action-dependent series (citylearn.py:1136-1323): control = the simulated values, baseline
                // = net without the storage devices' consumption (building.py:2886-2893); float32 values like the traced ones
                const double net = (double)(float)o.net;
                const double sto = (double)dyn_value<R>(CL_DYN_COOLING_STORAGE_ELECTRICITY_CONSUMPTION, c.p, s, o, (R)t_in)
                                 + (double)dyn_value<R>(CL_DYN_HEATING_STORAGE_ELECTRICITY_CONSUMPTION, c.p, s, o, (R)t_in)
                                 + (double)dyn_value<R>(CL_DYN_DHW_STORAGE_ELECTRICITY_CONSUMPTION, c.p, s, o, (R)t_in)
                                 + (double)dyn_value<R>(CL_DYN_ELECTRICAL_STORAGE_ELECTRICITY_CONSUMPTION, c.p, s, o, (R)t_in);
                const double nws = net - sto;
                const double price = (double)row[c.c_price], carbon = (double)row[c.c_carbon];
                double* a = s_kacc + tid;
                a[CL_KPI_EC * nt] += fmax(net, 0.0); a[CL_KPI_ZNE * nt] += net;
                a[CL_KPI_EMISSION * nt] += fmax((double)(float)o.emission, 0.0); a[CL_KPI_COST * nt] += fmax((double)(float)o.cost, 0.0);
                a[CL_KPI_B_EC * nt] += fmax(nws, 0.0); a[CL_KPI_B_ZNE * nt] += nws;
                a[CL_KPI_B_EMISSION * nt] += fmax(carbon * nws, 0.0); a[CL_KPI_B_COST * nt] += fmax(price * nws, 0.0);
                s_knws[pb * nt + ul] = nws;
            }
            if (fresh_tab) {
                // this unit's action-dependent observation columns go into its env's row image (loaded one step ago)
                mbar_wait(s_bar + 3 + pb, (uint32_t)((k >> 1) & 1));
                float* img = smf + lo.tmpl + ((size_t)pb * n_slots + e_l) * lo.Lp - k0;
                for (int i = dyn_lo; i < dyn_hi; ++i) {
                    const int2 cs = __ldg(d.dyn_cols + i);
                    float v = dyn_value<R>(cs.y, c.p, s, o, (R)t_in);
                    if (d.obs_t) v = transform_obs(d.obs_t, cs.x, v);
                    img[cs.x] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the TMA stores
            }
#ifdef CL_PHASE_TIMING
            if (want_dyn) {
#else
            if (trace != nullptr || want_dyn) {
#endif
                float dyn[CL_NDYN];
                fill_dyn<R>(c.p, s, o, (R)t_in, dyn);
#ifndef CL_PHASE_TIMING
                if (trace != nullptr) {
#pragma unroll
                    for (int j = 0; j < CL_NDYN; ++j) trace[(size_t)u * CL_NDYN + j] = dyn[j];
                }
#endif
                if (want_dyn) {
#pragma unroll
                    for (int j = 0; j < CL_NDYN; ++j) s_dynbuf[ul * CL_NDYN + j] = dyn[j];
                }
            }
        }
        CL_STAMP(4);
        if (WIDE) {
            // the tile's partial district sums: fixed-order shuffle tree per warp, one slot per warp; the cluster barrier below
            // publishes them to every CTA of the env (distributed shared memory)
            float vn = active ? (float)o.net : 0.f, vc = active ? (float)o.cost : 0.f, ve = active ? (float)o.emission : 0.f;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                vn += __shfl_xor_sync(0xffffffffu, vn, m); vc += __shfl_xor_sync(0xffffffffu, vc, m); ve += __shfl_xor_sync(0xffffffffu, ve, m);
            }
            if (lane == 0) {
                const int w = tid >> 5;
                s_wpart[(pb * 3 + 0) * 32 + w] = vn; s_wpart[(pb * 3 + 1) * 32 + w] = vc; s_wpart[(pb * 3 + 2) * 32 + w] = ve;
            }
            if (coupled) cluster_sync_all(); else __syncthreads();             // S1 (for the whole cluster when tiles exchange sums)
        } else if (dec) {
            __syncwarp();
            if (lane == 0) mbar_arrive(s_bar + 7 + pb);                        // R: done with red[pb], row t, the inputs of parity pb
        } else {
            __syncthreads();                                                   // S1: red / dynbuf / next-step building inputs complete
        }
        CL_STAMP(5);
        if (!dec && uniform && tid == 0 && k + 3 <= K) {
            // nobody reads row t any more (reward inputs were captured above): its slot takes the row of step t + 3
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(s_bar + slot_t, row_bytes);
            tma_load_1d(s_rows + slot_t * Wp, d.table + (size_t)(d.start0 + t + 3) * Wp, row_bytes, s_bar + slot_t);
        }
        // district sums in building order, like the reference's sum() over buildings (citylearn.py:1908-1918);
        // one thread per (quantity, env)
        if (WIDE) {
            const int nw = np_ >> 5;
            if (coupled) {
                // district sums = partial sums of every (tile, warp) of the cluster, read through distributed shared memory by
                // warp 0 of every CTA: one slot per lane, then a fixed-order shuffle tree (float32 like the reference, but a
                // different association than its left-to-right sum(): DESIGN.md 'Wide districts')
                if (tid < 32) {
                    const int total = NT * nw;
#pragma unroll 1
                    for (int q = 0; q < 3; ++q) {
                        if (q > 0 && district == nullptr) break;
                        const uint32_t local = smem_u32(s_wpart + (pb * 3 + q) * 32);
                        float acc = 0.f;
                        for (int idx = lane; idx < total; idx += 32) {
                            const int r = idx / nw, w = idx - r * nw;
                            acc += ld_cluster_f32(cluster_map(local, (uint32_t)r) + 4u * (uint32_t)w);
                        }
#pragma unroll
                        for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
                        if (lane == 0) {
                            if (q == 0) s_dsum[pb * epb] = acc;
                            if (district != nullptr && rank == 0) district[((size_t)k * d.E + e0) * 3 + q] = acc;
                        }
                    }
                }
            } else if (district != nullptr && tid < 3) {
                // independent tiles: this tile's partial sums go to the scratch buffer [K][E][tiles][3]; district_finish_kernel
                // adds the tiles up in order after the launch
                const float* src = s_wpart + (pb * 3 + tid) * 32;
                float acc = 0.f;
                for (int w = 0; w < nw; ++w) acc += src[w];
                district[(((size_t)k * d.E + e0) * NT + rank) * 3 + tid] = acc;
            }
        } else if (dec) {
            // the sums of step k - 1 (and, on the last step of the launch, of step k as well): its arrivals are a whole step old
            for (int kk = (k > 0 ? k - 1 : k); kk <= k; ++kk) {
                if (kk == k && k + 1 < K) break;
                const int pp = kk & 1;
                mbar_wait(s_bar + 7 + pp, (uint32_t)((kk >> 1) & 1));          // R: every warp's results of step kk are in red[pp]
                if (district != nullptr) {
                    const float* redk = smf + lo.red + pp * 3 * nt;
                    for (int idx = tid; idx < 3 * n_env; idx += np_) {
                        const int q = idx / n_env, le = idx - q * n_env;
                        const float* src = redk + q * nt + le * B;
                        float acc = 0.f;
                        int j = 0;
                        for (; j + 4 <= B; j += 4) { acc += src[j]; acc += src[j + 1]; acc += src[j + 2]; acc += src[j + 3]; }   // same left-to-right order
                        for (; j < B; ++j) acc += src[j];
                        district[((size_t)kk * d.E + e0 + le) * 3 + q] = acc;
                    }
                }
                __syncwarp();
                if (lane == 0 && kk + 2 < K) mbar_arrive(s_bar + 9 + pp);      // S: red[pp] may be written again (step kk + 2)
            }
        } else
        for (int idx = tid; idx < 3 * n_env; idx += np_) {
            const int q = idx / n_env, le = idx - q * n_env;
            if (q > 0 && district == nullptr) continue;
            const float* src = red + q * nt + le * B;
            float acc = 0.f;
            int j = 0;
            for (; j + 4 <= B; j += 4) { acc += src[j]; acc += src[j + 1]; acc += src[j + 2]; acc += src[j + 3]; }   // same left-to-right order
            for (; j < B; ++j) acc += src[j];
            // building-sharded district: `acc` covers this rank's buildings; complete it over the ranks through peer memory
            if (d.x_n > 1) acc = exchange_sum(d, acc, e0 + le, q, d.x_epoch + (unsigned)k + 1u);
            if (q == 0) s_dsum[pb * epb + le] = acc;
            if (district != nullptr) district[((size_t)k * d.E + e0 + le) * 3 + q] = acc;
            if (kpi && q == 0) {
                // district series of this env: control = the district net, baseline = sum over buildings of the net without storage
                const double* nw = s_knws + pb * nt + le * B;
                double tot = 0.0;
                for (int jj = 0; jj < B; ++jj) tot += nw[jj];
                kpi_push(s_kenv + (size_t)(le * 2 + 0) * CL_NKPI_ENV, (double)acc);
                kpi_push(s_kenv + (size_t)(le * 2 + 1) * CL_NKPI_ENV, tot);
            }
        }
        if (has_ev && d.ev_n > 0) {
            // next_time_step for the vehicles (citylearn.py:1336-1351): the soc[t+1] entry starts at 0, then the row's compiled operation
            // (ev.compile_schedule): drift of an away vehicle, pre-connection SOC, arrival SOC of a new connection.  Every charger of the
            // block has finished step t (S1); the barrier at the end of the iteration publishes the entries to step t + 1.
            const size_t EV = (size_t)d.E * d.ev_n;
            const int tn = t + 1;
            for (int idx = tid; idx < n_env * d.ev_n; idx += np_) {
                const int le = idx / d.ev_n, v = idx - le * d.ev_n;
                const size_t i = (size_t)(e0 + le) * d.ev_n + v;
                const float prev = d.ev_sf[EV + i];
                float cur = 0.f;
                if (tn <= d.T - 1) {
                    const int4 ec = __ldg(reinterpret_cast<const int4*>(d.ev_cols) + v);
                    const float* rn = uniform ? row_next : d.table + (size_t)(start_e + tn) * Wp;
                    if (tn + 1 < d.T) {                                   // simulate_unconnected_ev_soc returns early on the last step
                        const double dr = __ldg(d.ev_drift + (size_t)(d.start0 + tn) * d.ev_n + v);
                        if (dr == dr) cur = (float)fmin(fmax((double)prev * dr, 0.0), 1.0);
                        const float sm = rn[ec.y];
                        if (sm == sm) cur = sm;
                    }
                    const float as = rn[ec.x];
                    if (as == as) cur = as;
                }
                d.ev_sf[i] = prev; d.ev_sf[EV + i] = cur;
            }
        }
        if (need_dsum) __syncthreads();                                        // S2 only when a reward reads the district sum
        if (fused_reward) {
            float r = 0.f;
            if (active) {
                if (need_dsum) ri.district_net = s_dsum[pb * epb + e_l];
                if constexpr (has_ev) r = d.reward_id == CL_REWARD_ELECTRIC_VEHICLES ? ev_reward(chi, n_chi, ri.net, ri.district_net, t, cc_penalty) : unit_reward(d.reward_id, d.rp, ri);
                else r = unit_reward(d.reward_id, d.rp, ri);
            }
            float* rk = reward + (size_t)k * d.E * Rdim;
            if (d.central && WIDE) {
                double v = active ? (double)r : 0.0;
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                if (lane == 0) s_rpart[pb * 32 + (tid >> 5)] = v;
                cluster_sync_all();
                if (rank == 0 && tid == 0) {
                    const uint32_t local = smem_u32(s_rpart + pb * 32);
                    double acc = 0.0;
                    for (int rr = 0; rr < NT; ++rr) {
                        const uint32_t ra = cluster_map(local, (uint32_t)rr);
                        for (int w = 0; w < (np_ >> 5); ++w) acc += ld_cluster_f64(ra + 8u * (uint32_t)w);
                    }
                    rk[e0] = (float)acc;
                }
            } else if (d.central) {
                if (k > 0) __syncthreads();                                    // previous step's rsum readers are done
                s_rsum[ul] = r;
                __syncthreads();
                if (tid < n_env) {
                    float sr = 0.f;
                    if (d.reward_id == CL_REWARD_MARL || d.reward_id == CL_REWARD_SOLAR_PENALTY_AND_COMFORT) {
                        double acc = 0.0;   // these rewards are float64 in the reference
                        for (int j = 0; j < B; ++j) acc += (double)s_rsum[tid * B + j];
                        sr = (float)acc;
                    } else {
                        for (int j = 0; j < B; ++j) sr += s_rsum[tid * B + j];
                    }
                    rk[e0 + tid] = sr;
                }
            } else if (active) {
                rk[u] = r;
            }
        }
        CL_STAMP(6);
        if (obs != nullptr && !tmpl_path && !fresh_tab) {
            write_obs_general(d, obs + (size_t)k * d.E * d.L, e0, n_env, t + 1, want_dyn ? s_dynbuf : nullptr, tid, np_, k0, k1, b0, nb);
            if (want_dyn && k + 1 < K) __syncthreads();                         // dynbuf is single-buffered
        }
        if (has_ev && k + 1 < K) __syncthreads();                               // the vehicles' soc[t+1] entries are in place for step t + 1
    }
#ifdef CL_PHASE_TIMING
    { const int k = K - 1; CL_STAMP(7); }
#endif
    if (kpi) {
        __syncthreads();                                 // the last step's env-level pushes are complete
        if (active && !is_helper) {
#pragma unroll
            for (int j = 0; j < CL_NKPI_UNIT; ++j) d.kpi_unit[(size_t)u * CL_NKPI_UNIT + j] += s_kacc[j * nt + tid];
        }
        for (int i = tid; i < n_env * 2 * CL_NKPI_ENV; i += nt) d.kpi_env[(size_t)e0 * 2 * CL_NKPI_ENV + i] = s_kenv[i];
    }
    if (is_helper) tma_store_wait_all<0>();
    if (active && !is_helper) store_state<R, THERMAL>(d, u, s);
    if (coupled) cluster_sync_all();  // nobody exits while a peer may still read its partial sums
}

// ------------------------------------------------------------------------------------------------------------------
// online KPI accumulators (cl_kpi_*): one block per env, threads stride over the buildings
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void kpi_push(double* a, double x) {
    const double n = a[CL_KE_N];
    if (n > 0.0) a[CL_KE_RAMP] += fmax(x - a[CL_KE_PREV], 0.0);          // CostFunction.ramping: positive ramps only
    a[CL_KE_PREV] = x;
    a[CL_KE_ALL_MAX] = n > 0.0 ? fmax(a[CL_KE_ALL_MAX], x) : x;
    a[CL_KE_N] = n + 1.0;
    a[CL_KE_D_SUM] += x; a[CL_KE_D_MAX] = a[CL_KE_D_CNT] > 0.0 ? fmax(a[CL_KE_D_MAX], x) : x; a[CL_KE_D_CNT] += 1.0;
    if (a[CL_KE_D_CNT] == 24.0) {
        a[CL_KE_D_FIN_LF] += 1.0 - (a[CL_KE_D_SUM] / 24.0) / a[CL_KE_D_MAX];
        a[CL_KE_D_FIN_PEAK] += a[CL_KE_D_MAX]; a[CL_KE_D_FIN_N] += 1.0;
        a[CL_KE_D_SUM] = 0.0; a[CL_KE_D_CNT] = 0.0;
    }
    a[CL_KE_M_SUM] += x; a[CL_KE_M_MAX] = a[CL_KE_M_CNT] > 0.0 ? fmax(a[CL_KE_M_MAX], x) : x; a[CL_KE_M_CNT] += 1.0;
    if (a[CL_KE_M_CNT] == 730.0) {
        a[CL_KE_M_FIN_LF] += 1.0 - (a[CL_KE_M_SUM] / 730.0) / a[CL_KE_M_MAX];
        a[CL_KE_M_FIN_N] += 1.0;
        a[CL_KE_M_SUM] = 0.0; a[CL_KE_M_CNT] = 0.0;
    }
}
__global__ void kpi_accumulate_kernel(Dev d, int t, const float* __restrict__ trace, const float* __restrict__ district,
                                      double* __restrict__ ku, double* __restrict__ ke) {
    __shared__ double s_part[32];
    const int e = blockIdx.x;
    const float* row = d.table + (size_t)(__ldg(d.start + e) + t) * d.Wp;
    double base_sum = 0.0;
    for (int b = threadIdx.x; b < d.B; b += blockDim.x) {
        const float* tr = trace + ((size_t)e * d.B + b) * CL_NDYN;
        const double net = tr[CL_DYN_NET_ELECTRICITY_CONSUMPTION];
        const double sto = (double)tr[CL_DYN_COOLING_STORAGE_ELECTRICITY_CONSUMPTION] + (double)tr[CL_DYN_HEATING_STORAGE_ELECTRICITY_CONSUMPTION]
                         + (double)tr[CL_DYN_DHW_STORAGE_ELECTRICITY_CONSUMPTION] + (double)tr[CL_DYN_ELECTRICAL_STORAGE_ELECTRICITY_CONSUMPTION];
        const double nws = net - sto;                                     // net_electricity_consumption_without_storage (building.py:2886-2893)
        const double price = row[__ldg(d.ip + CL_IP_C_PRICE * d.B + b)], carbon = row[__ldg(d.ip + CL_IP_C_CARBON * d.B + b)];
        double* a = ku + ((size_t)e * d.B + b) * CL_NKPI_UNIT;
        a[CL_KPI_EC] += fmax(net, 0.0); a[CL_KPI_ZNE] += net;
        a[CL_KPI_EMISSION] += fmax((double)tr[CL_DYN_NET_ELECTRICITY_CONSUMPTION_EMISSION], 0.0);
        a[CL_KPI_COST] += fmax((double)tr[CL_DYN_NET_ELECTRICITY_CONSUMPTION_COST], 0.0);
        a[CL_KPI_B_EC] += fmax(nws, 0.0); a[CL_KPI_B_ZNE] += nws;
        a[CL_KPI_B_EMISSION] += fmax(carbon * nws, 0.0); a[CL_KPI_B_COST] += fmax(price * nws, 0.0);
        base_sum += nws;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) base_sum += __shfl_xor_sync(0xffffffffu, base_sum, m);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = base_sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) tot += s_part[w];
        kpi_push(ke + ((size_t)e * 2 + 0) * CL_NKPI_ENV, (double)district[(size_t)e * 3]);
        kpi_push(ke + ((size_t)e * 2 + 1) * CL_NKPI_ENV, tot);
    }
}

// district sums of independent tiles: district[n][q] = sum over tiles (in order) of part[n][tile][q], n = (step, env)
__global__ void district_finish_kernel(const float* __restrict__ part, float* __restrict__ district, long n3, int tiles) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const long n = i / 3; const int q = (int)(i - n * 3);
    float acc = 0.f;
    for (int r = 0; r < tiles; ++r) acc += part[(n * tiles + r) * 3 + q];
    district[i] = acc;
}

// ------------------------------------------------------------------------------------------------------------------
// observation table: obs_tab[r][k] = transformed observation column k at the time step whose table row is r (stale DYN columns
// and outage columns are 0 here; outage columns are patched per step)
// ------------------------------------------------------------------------------------------------------------------
__global__ void build_obs_table_kernel(Dev d, float* __restrict__ out) {
    for (int r = blockIdx.y; r < d.n_rows; r += gridDim.y)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < d.obs_pitch; k += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (k < d.L) {
            const int cc = __ldg(d.tcol + k);
            if (cc >= 0) v = __ldg(d.table + (size_t)r * d.Wp + cc);
            if (d.obs_t) v = transform_obs(d.obs_t, k, v);
        }
        out[(size_t)r * d.obs_pitch + k] = v;
    }
}

// the env-independent observation rows of time steps t_first .. t_first + n - 1 (reference-parity observations after a step, SURVEY
// A.6-1): what every env's row holds - for callers that want ONE row per step instead of E identical ones (cl_obs_rows)
__global__ void obs_rows_kernel(Dev d, int t_first, int n, float* __restrict__ out) {
    const long total = (long)n * d.L;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int r = (int)(i / d.L), k = (int)(i - (long)r * d.L);
        const int t_obs = t_first + r;
        const int cc = __ldg(d.tcol + k);
        float v = 0.f;
        if (cc >= 0) v = __ldg(d.table + (size_t)(d.start0 + t_obs) * d.Wp + cc);
        else if (cc <= -2 && d.has_outage) v = __ldg(d.outage + (-2 - cc) * d.T + t_obs);
        if (d.obs_t) v = transform_obs(d.obs_t, k, v);
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// reset kernel: state <- initial values; obs <- observation at t = 0 (citylearn.py:1829-1886, building.py:2526-2564)
// ------------------------------------------------------------------------------------------------------------------
template <typename R, bool THERMAL, int MAXT>
__global__ void __launch_bounds__(MAXT) reset_kernel(Dev d, float* __restrict__ obs) {
    extern __shared__ __align__(16) float smf[];
    const int nt = blockDim.x, tid = threadIdx.x;
    const int B = d.B, epb = d.envs_per_block;
    const SmemLayout lo = smem_layout(d.tile_b, d.Wp, d.Lt, epb, nt, (int)sizeof(R), d.lstm_smem, d.tab_layout, d.fresh_slots, d.ev_n > 0 ? d.tile_b + d.ev_n : d.n_curves, d.kpi_smem);
    float* s_dynbuf = smf + lo.dynbuf;
    // building tiles (wide districts): block (group, rank) owns buildings [b0, b0 + nb) and observation columns [k0, k1)
    const int rank = (int)blockIdx.x % d.tiles;
    const int b0 = rank * d.tile_b, nb = min(d.tile_b, B - b0);
    const int k0 = d.tiles > 1 ? __ldg(d.tile_k + rank) : 0, k1 = d.tiles > 1 ? __ldg(d.tile_k + rank + 1) : d.L;
    const int e0 = ((int)blockIdx.x / d.tiles) * epb;
    const int n_env = min(epb, d.E - e0);
    const int n_units = n_env * nb;
    const int e_l = tid / nb, b = b0 + (tid - e_l * nb);
    const int e = e0 + e_l, u = e * B + b;
    if (tid < n_units) {
        const auto* P = PSel<R>::p(d);
        UnitCtx<R> c;
        load_ctx<R, THERMAL>(d, b, c);
        UnitState<R> s;
        s.soc_b = Num<R>::r32((R)__ldg(P + CL_P_BAT_INITIAL_SOC * B + b));
        s.cap_deg = c.p.bat_capacity;
        s.rte_b = Num<R>::sqrt_((R)__ldg(P + CL_P_BAT_EFFICIENCY0 * B + b));
        s.soc_cs = Num<R>::r32((R)__ldg(P + CL_P_CS_INITIAL_SOC * B + b));
        s.soc_hs = Num<R>::r32((R)__ldg(P + CL_P_HS_INITIAL_SOC * B + b));
        s.soc_ds = Num<R>::r32((R)__ldg(P + CL_P_DS_INITIAL_SOC * B + b));
        store_state<R, true>(d, u, s);
        if (d.lst != nullptr) {
            for (int j = 0; j < kLstmStateFloats; ++j) d.lst[(size_t)j * d.U + u] = 0.f;   // dynamics.py:112-127
        }
        if (obs != nullptr) {
            const RowRef row = {d.table + (size_t)__ldg(d.start + e) * d.Wp, 0u};
            UnitInputs<R> in;
            load_inputs<R, THERMAL>(d, c, row, b, 0, in);
            UnitResult<R> o;
            unit_time0<R, THERMAL>(c.p, in, o);
            fill_dyn<R>(c.p, s, o, (R)row[c.c_tin], s_dynbuf + tid * CL_NDYN);
        }
    }
    if ((d.ev_n + d.wm_n) > 0 && rank == 0) {
        // ElectricVehicle.reset + associate_chargers_to_electric_vehicles at t = 0 (citylearn.py:1871-1874): soc[0] = initial SOC, replaced
        // by the arrival SOC of a vehicle that is plugged in on the episode's first row
        const size_t EV = (size_t)d.E * d.ev_n;
        for (int idx = tid; idx < n_env * d.ev_n; idx += nt) {
            const int le = idx / d.ev_n, v = idx - le * d.ev_n;
            const size_t i = (size_t)(e0 + le) * d.ev_n + v;
            const double* ep = d.ev_pd + (size_t)v * CL_NPARAM;
            const float t0v = __ldg(d.table + (size_t)__ldg(d.start + e0 + le) * d.Wp + __ldg(d.ev_cols + v * 4 + 3));
            d.ev_sf[i] = 0.f;
            d.ev_sf[EV + i] = t0v == t0v ? t0v : (float)__ldg(ep + CL_P_BAT_INITIAL_SOC);
            d.ev_sd[i] = __ldg(ep + CL_P_BAT_CAPACITY);
            d.ev_sd[EV + i] = sqrt(__ldg(ep + CL_P_BAT_EFFICIENCY0));
            d.ev_flag[i] = 0;
        }
        for (int idx = tid; idx < n_env * d.wm_n; idx += nt) d.wm_flag[(size_t)e0 * d.wm_n + idx] = 0;
    }
    if (obs != nullptr) {
        __syncthreads();
        write_obs_general(d, obs, e0, n_env, 0, s_dynbuf, tid, nt, k0, k1, b0, nb);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                                        \
    do {                                                                                                   \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) return fail(CL_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)

}  // namespace cl

struct cl_env {
    cl::Dev d;
    int precision = 0;
    bool thermal = false;
    int t = -1;              // -1: not reset
    int32_t* t_dev = nullptr;       // device-resident copy of t (cl_device_time_enable / cl_advance_device)
    bool device_time = false;       // sticky: the device counter is authoritative (graph replays advance it without the host)
    // building-sharded district (cl_exchange_*)
    uint2* x_buf = nullptr;          // own slot array [2][n][E][3]
    uint2** x_peers_dev = nullptr;   // device array of the ranks' slot arrays as mapped here
    int32_t* x_err_dev = nullptr;
    std::vector<void*> x_opened;     // cudaIpcOpenMemHandle mappings to close
    unsigned x_epoch = 0;            // steps exchanged so far
    size_t ev_sf_floats = 0, ev_sd_doubles = 0, ev_flag_bytes = 0, wm_flag_bytes = 0;   // vehicle / washing-machine state (districts with cl_ev_desc)
    std::vector<float> lstm_packed;  // host copy of the packed LSTM weights when they fit the constant bank (<= kLstmConstBuildings buildings)
    int n_sm = 148;
    int T = 0;
    int threads = 0, blocks = 0;
    float* obs_tab_dev = nullptr;   // precomputed observation table (build_obs_table)
    double* kpi_unit = nullptr;     // online KPI accumulators (cl_kpi_enable)
    double* kpi_env = nullptr;
    volatile int32_t* host_flag = nullptr;   // page-locked completion flag of cl_step_host (polled by the host)
    int32_t host_seq = 0;
    bool decouple = false;          // CL_B200_DECOUPLE at cl_create: split-phase step barrier (advance_kernel<..., DEC>)
    bool kpi_fused = false;         // accumulated inside advance_kernel (else: cl_kpi_accumulate on the step's trace)
    float* dpart = nullptr;         // wide districts: per-tile partial district sums of one launch chunk
    size_t dpart_floats = 0;
    bool wide = false;       // building-tiled district: cluster launch of advance_kernel<..., WIDE = true>
    int64_t launches = 0;
    std::vector<void*> allocs;
    float* outage_dev = nullptr;
    int outage_T = 0;
    size_t st_floats = 0, dst_doubles = 0, lst_floats = 0;
    bool dynamics = false;
};

using namespace cl;

static cl_env* g_lstm_const_owner[64] = {};     // per device: the handle whose LSTM weights are in the constant bank (claim_lstm_constants)

template <typename T> static int dev_copy(cl_env* env, const T* host, size_t n, T** out) {
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    env->allocs.push_back(p);
    if (n) CUDA_TRY(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<T*>(p);
    return CL_OK;
}

// (Re)build the precomputed observation table of a reference-parity (stale-observation) district.  Skipped - the kernels then
// gather the row every step - when it would exceed CL_B200_OBS_TABLE_MB (default 4096 MiB) or the observation row is not a
// multiple of 16 bytes.
static bool obs_table_fits(const Dev& d) {
    if ((d.L & 3) != 0 || d.obs_state) return false;     // per-env state columns cannot come from a table shared by all envs
    long budget_mb = 4096;
    if (const char* ev = std::getenv("CL_B200_OBS_TABLE_MB")) budget_mb = std::atol(ev);
    return (size_t)d.n_rows * (size_t)((d.L + 3) & ~3) * sizeof(float) <= (size_t)budget_mb * 1024 * 1024;
}
static int build_obs_table(cl_env* env) {
    Dev& d = env->d;
    d.obs_tab = nullptr;
    if (!d.tab_layout) return CL_OK;               // decided with the launch geometry (cl_create)
    const int pitch = (d.L + 3) & ~3;
    const size_t bytes = (size_t)d.n_rows * pitch * sizeof(float);
    if (!env->obs_tab_dev) {
        if (cudaMalloc(reinterpret_cast<void**>(&env->obs_tab_dev), bytes) != cudaSuccess) {
            cudaGetLastError();
            env->obs_tab_dev = nullptr;
            return fail(CL_ERR_CUDA, "cl_create: not enough device memory for the observation table (lower CL_B200_OBS_TABLE_MB to use the gather path)");
        }
        env->allocs.push_back(env->obs_tab_dev);
    }
    d.obs_pitch = pitch;
    const int threads = 256;
    int bx = (pitch + threads - 1) / threads;
    if (bx > 64) bx = 64;
    build_obs_table_kernel<<<dim3((unsigned)bx, (unsigned)std::min(d.n_rows, 65535)), threads>>>(d, env->obs_tab_dev);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    d.obs_tab = env->obs_tab_dev;
    return CL_OK;
}

// opt in to large dynamic shared memory (both kernels, all instantiations).  The attribute is per function and process-wide: never
// lower it below what an earlier handle needs
static void ensure_smem_optin(size_t smem) {
    static size_t optin_max = 0;
    if (smem > optin_max) optin_max = smem;
// (the carve-out hint: prefer shared memory over L1 so that as many blocks as the registers allow are resident - the driver's default
    // carve-out fits ONE block of a kernel with large dynamic shared memory)
#define OPTIN(K) (cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin_max), cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared))
#define OPTIN4(K, M) OPTIN((K<float, false, M>)); OPTIN((K<float, true, M>)); OPTIN((K<double, false, M>)); OPTIN((K<double, true, M>))
#define OPTINA(M) OPTIN((advance_kernel<float, false, false, M>)); OPTIN((advance_kernel<float, true, false, M>)); OPTIN((advance_kernel<float, true, true, M>)); \
    OPTIN((advance_kernel<double, false, false, M>)); OPTIN((advance_kernel<double, true, false, M>)); OPTIN((advance_kernel<double, true, true, M>))
#define OPTINK(M) OPTIN((advance_kernel<float, false, false, M, false, true>)); OPTIN((advance_kernel<float, true, false, M, false, true>)); \
    OPTIN((advance_kernel<double, false, false, M, false, true>)); OPTIN((advance_kernel<double, true, false, M, false, true>))
    OPTINK(512); OPTINK(1024);
#undef OPTINK
#define OPTINE(M) OPTIN((advance_kernel<float, false, false, M, false, false, true>)); OPTIN((advance_kernel<float, true, false, M, false, false, true>)); \
    OPTIN((advance_kernel<double, false, false, M, false, false, true>)); OPTIN((advance_kernel<double, true, false, M, false, false, true>))
    OPTINE(512); OPTINE(1024);
#undef OPTINE
#define OPTIND(M) OPTIN((advance_kernel<float, false, false, M, false, false, false, true>)); OPTIN((advance_kernel<float, true, false, M, false, false, false, true>)); \
    OPTIN((advance_kernel<double, false, false, M, false, false, false, true>)); OPTIN((advance_kernel<double, true, false, M, false, false, false, true>))
    OPTIND(512); OPTIND(1024);
#undef OPTIND
    OPTINA(512); OPTINA(1024); OPTIN((advance_kernel<float, true, true, kDynMaxT>)); OPTIN((advance_kernel<double, true, true, kDynMaxT>)); OPTIN4(reset_kernel, 512); OPTIN4(reset_kernel, 1024);
    OPTIN((advance_kernel<float, false, false, 512, true>)); OPTIN((advance_kernel<float, true, false, 512, true>));
    OPTIN((advance_kernel<double, false, false, 512, true>)); OPTIN((advance_kernel<double, true, false, 512, true>));
#undef OPTINA
#undef OPTIN4
#undef OPTIN
}

extern "C" int cl_abi_version(void) { return CL_ABI_VERSION; }
extern "C" const char* cl_last_error(void) { return g_err.c_str(); }

extern "C" int cl_create(const cl_district_desc* desc, cl_env** out) {
    if (!desc || !out) return fail(CL_ERR_INVALID, "cl_create: null argument");
    if (desc->abi_version != CL_ABI_VERSION) return fail(CL_ERR_INVALID, "cl_create: ABI version mismatch");
    if (desc->n_buildings < 1 || desc->n_envs < 1 || desc->n_rows < 2 || desc->n_cols < 1 || desc->obs_dim < 1)
        return fail(CL_ERR_INVALID, "cl_create: empty district");
    if (desc->n_buildings > 8 * 480) return fail(CL_ERR_UNSUPPORTED, "cl_create: more than 3840 buildings per district (8 tiles of 480) not supported");
    if (!desc->table || !desc->params || !desc->iparams || !desc->obs_desc) return fail(CL_ERR_INVALID, "cl_create: null table/params");
    if (desc->precision != CL_PRECISION_FP32 && desc->precision != CL_PRECISION_FP64) return fail(CL_ERR_INVALID, "cl_create: bad precision");
    if (desc->reward_id < CL_REWARD_NONE || desc->reward_id > CL_REWARD_ELECTRIC_VEHICLES) return fail(CL_ERR_INVALID, "cl_create: unknown reward_id");
    if (desc->reward_id == CL_REWARD_ELECTRIC_VEHICLES && !(desc->ev && desc->ev->n_chargers + desc->ev->n_machines > 0))
        return fail(CL_ERR_INVALID, "cl_create: CL_REWARD_ELECTRIC_VEHICLES needs a district with chargers (cl_district_desc.ev)");
    cl_env* env = new (std::nothrow) cl_env();
    if (!env) return fail(CL_ERR_INVALID, "cl_create: out of host memory");
    { const char* v = std::getenv("CL_B200_DECOUPLE"); env->decouple = v != nullptr && v[0] != '\0' && v[0] != '0'; }
    Dev& d = env->d;
    std::memset(&d, 0, sizeof(d));
    const int B = desc->n_buildings;
    d.B = B; d.E = desc->n_envs; d.U = d.B * d.E; d.n_rows = desc->n_rows; d.W = desc->n_cols; d.Wp = (desc->n_cols + 3) & ~3;
    d.A = desc->action_dim; d.L = desc->obs_dim; d.central = desc->central_agent; d.reward_id = desc->reward_id;
    d.stale = desc->stale_observations;
    for (int i = 0; i < 8; ++i) d.rp[i] = (float)desc->reward_params[i];
    env->precision = desc->precision;
    int any_thermal = 0, any_dyn = 0, nmax = 2;
    for (int b = 0; b < B; ++b) {
        const int f = desc->iparams[CL_IP_FLAGS * B + b];
        any_thermal |= (f & CL_F_HAS_THERMAL);
        any_dyn |= (f & CL_F_DYNAMICS);
        const int pe = desc->iparams[CL_IP_PE_N * B + b], cp = desc->iparams[CL_IP_CP_N * B + b];
        if (pe < 2 || pe > CL_MAX_CURVE || cp < 2 || cp > CL_MAX_CURVE) { delete env; return fail(CL_ERR_INVALID, "cl_create: battery curves need 2 .. 8 points"); }
        nmax = std::max(nmax, std::max(pe, cp));
        // the segment search counts the points below x: the curve abscissae must ascend (energy_model.py:1083-1109 assumes it too)
        for (int w = 0; w < 2; ++w)
            for (int k = 0; k + 1 < (w ? cp : pe); ++k) {
                const double* xs = desc->params + (size_t)(w ? CL_P_CP_X0 : CL_P_PE_X0) * B;
                if (!(xs[(size_t)k * B + b] <= xs[(size_t)(k + 1) * B + b]) || !(xs[(size_t)k * B + b] >= 0.0))
                    { delete env; return fail(CL_ERR_INVALID, "cl_create: battery curve abscissae must be non-negative and ascending"); }
            }
    }
    d.curve_nmax = nmax;
    {
        // uniform-grid index of the curve abscissae (unit_physics.cuh, SmemCurves): usable when no cell holds two points of a curve.
        // The cells are those of the values the kernel compares, i.e. of the float-rounded abscissae in CL_PRECISION_FP32.
        const int n_ev_l = desc->ev ? desc->ev->n_ev : 0;          // the vehicles' battery curves follow the buildings'
        std::vector<uint8_t> lut((size_t)(B + n_ev_l) * 2 * kCurveLutStride, 0);
        bool ok = std::getenv("CL_B200_NO_CURVE_LUT") == nullptr;
        for (int b = 0; b < B && ok; ++b)
            for (int w = 0; w < 2 && ok; ++w)
                ok = build_curve_lut(desc->params + (size_t)(w ? CL_P_CP_X0 : CL_P_PE_X0) * B + b, (size_t)B,
                                     desc->iparams[(w ? CL_IP_CP_N : CL_IP_PE_N) * B + b], desc->precision == CL_PRECISION_FP32,
                                     &lut[((size_t)b * 2 + w) * kCurveLutStride]);
        for (int v = 0; v < n_ev_l && ok; ++v)
            for (int w = 0; w < 2 && ok; ++w) {
                const int n = desc->ev->ev_iparams[v * 2 + w];
                if (n < 2 || n > CL_MAX_CURVE) { delete env; return fail(CL_ERR_INVALID, "cl_create: vehicle battery curves need 2 .. 8 points"); }
                d.curve_nmax = std::max(d.curve_nmax, n);
                ok = build_curve_lut(desc->ev->ev_params + (size_t)v * CL_NPARAM + (w ? CL_P_CP_X0 : CL_P_PE_X0), 1, n,
                                     desc->precision == CL_PRECISION_FP32, &lut[((size_t)(B + v) * 2 + w) * kCurveLutStride]);
            }
        if (ok) {
            uint8_t* p = nullptr;
            int rc = dev_copy(env, lut.data(), lut.size(), &p);
            if (rc) { cl_destroy(env); return rc; }
            d.curve_lut = p;
        }
    }
    {
        // distinct curve sets (point counts + the 32 abscissae / ordinates): shared-memory tables are per distinct set when there are few
        std::vector<int32_t> cid((size_t)B, 0), rep;
        for (int b = 0; b < B; ++b) {
            int found = -1;
            for (size_t c = 0; c < rep.size() && found < 0; ++c) {
                const int r = rep[c];
                bool same = desc->iparams[CL_IP_PE_N * B + b] == desc->iparams[CL_IP_PE_N * B + r] && desc->iparams[CL_IP_CP_N * B + b] == desc->iparams[CL_IP_CP_N * B + r];
                for (int k = 0; k < 4 * CL_MAX_CURVE && same; ++k)
                    same = std::memcmp(&desc->params[(size_t)(CL_P_PE_X0 + k) * B + b], &desc->params[(size_t)(CL_P_PE_X0 + k) * B + r], sizeof(double)) == 0;
                if (same) found = (int)c;
            }
            if (found < 0) { if (rep.size() >= 64) { rep.clear(); break; } found = (int)rep.size(); rep.push_back(b); }
            cid[(size_t)b] = found;
        }
        if (!rep.empty() && std::getenv("CL_B200_NO_CURVE_SHARING") == nullptr && !(desc->ev && desc->ev->n_ev > 0)) {
            int32_t *pc = nullptr, *pr = nullptr;
            int rc = dev_copy(env, cid.data(), cid.size(), &pc);
            if (!rc) rc = dev_copy(env, rep.data(), rep.size(), &pr);
            if (rc) { cl_destroy(env); return rc; }
            d.curve_id = pc; d.curve_rep = pr; d.n_curves = (int)rep.size();
        }
    }
    env->thermal = any_thermal != 0;
    d.any_dynamics = any_dyn != 0;
    env->dynamics = any_dyn != 0;
    if (any_dyn) {
        // repack every building's LSTM block (schema.py order: W_ih0 [64,nin], W_hh0 [64,16], b_ih0, b_hh0, W_ih1 [64,16], W_hh1, b_ih1,
        // b_hh1, w_lin [16], b_lin) into the padded device layout of unit_physics.cuh
        if (!desc->lstm_weights) { delete env; return fail(CL_ERR_INVALID, "cl_create: dynamics buildings without lstm_weights"); }
        std::vector<float> packed((size_t)B * kLstmStride, 0.f);
        for (int b = 0; b < B; ++b) {
            if (!(desc->iparams[CL_IP_FLAGS * B + b] & CL_F_DYNAMICS)) continue;
            const int nin = desc->iparams[CL_IP_DYN_N_INPUTS * B + b], H = desc->iparams[CL_IP_DYN_HIDDEN * B + b];
            const int L = desc->iparams[CL_IP_DYN_LOOKBACK * B + b], off = desc->iparams[CL_IP_DYN_W_OFFSET * B + b];
            if (H < 1 || H > kLstmH || nin < 1 || nin > kLstmIn || L < 1 || L > kLstmMaxLookback) {
                delete env;
                return fail(CL_ERR_UNSUPPORTED, "cl_create: LSTM dynamics supports hidden_size <= 16, <= 16 inputs, lookback <= 12, 2 layers");
            }
            // source blocks have their natural sizes ([4H, nin], [4H, H], [4H] ...); the device layout is padded to 16 hidden
            // units / 16 inputs.  Padded hidden units have zero weights and biases: their cell and output stay exactly 0.
            const int G = 4 * H;
            const size_t need = (size_t)G * nin + (size_t)G * H + 2 * G + 2 * (size_t)G * H + 2 * G + H + 1;
            if (off < 0 || (size_t)off + need > (size_t)desc->lstm_weight_count) { delete env; return fail(CL_ERR_INVALID, "cl_create: lstm_weights block out of range"); }
            const float* w = desc->lstm_weights + off;
            float* o = &packed[(size_t)b * kLstmStride];
            const float* wih0 = w; const float* whh0 = wih0 + G * nin; const float* bih0 = whh0 + G * H; const float* bhh0 = bih0 + G;
            const float* wih1 = bhh0 + G; const float* whh1 = wih1 + G * H; const float* bih1 = whh1 + G * H; const float* bhh1 = bih1 + G;
            const float* wl = bhh1 + G; const float* bl = wl + H;
            for (int q = 0; q < 4; ++q) {
                for (int j = 0; j < H; ++j) {
                    const int rs = q * H + j, r = q * kLstmH + j;      // gate-major rows (i, f, g, o)
                    for (int i = 0; i < nin; ++i) o[r * 16 + i] = wih0[rs * nin + i];
                    for (int i = 0; i < H; ++i) o[64 * 16 + r * 16 + i] = whh0[rs * H + i];
                    o[64 * 32 + r] = bih0[rs] + bhh0[rs];
                    for (int i = 0; i < H; ++i) o[kLstmLayerStride + r * 16 + i] = wih1[rs * H + i];
                    for (int i = 0; i < H; ++i) o[kLstmLayerStride + 64 * 16 + r * 16 + i] = whh1[rs * H + i];
                    o[kLstmLayerStride + 64 * 32 + r] = bih1[rs] + bhh1[rs];
                }
            }
            for (int i = 0; i < H; ++i) o[2 * kLstmLayerStride + i] = wl[i];
            o[2 * kLstmLayerStride + 16] = bl[0];
        }
        float* pw = nullptr;
        int rc = dev_copy(env, packed.data(), packed.size(), &pw);
        if (rc) { cl_destroy(env); return rc; }
        d.lstm_w = pw;
#ifdef CL_LSTM_CONST
        if (B <= kLstmConstBuildings && std::getenv("CL_B200_NO_LSTM_CONST") == nullptr) env->lstm_packed = packed;
#endif
        d.lstm_smem = ((size_t)B * kLstmStride * sizeof(float) <= 120 * 1024) ? 1 : 0;
        // + the tensor-core operand fragments when they fit next to everything else (3 LSTM buildings: 134 KB in all)
        if (d.lstm_smem && (size_t)B * (kLstmStride + kLstmFragFloats + kLstmPreRing * 64) * sizeof(float) <= 150 * 1024 &&
            std::getenv("CL_B200_NO_LSTM_MMA") == nullptr) d.lstm_smem = 2;
    }
    // padded table
    {
        std::vector<float> padded((size_t)d.n_rows * d.Wp, 0.f);
        for (int r = 0; r < d.n_rows; ++r) std::memcpy(&padded[(size_t)r * d.Wp], desc->table + (size_t)r * d.W, sizeof(float) * d.W);
        float* p = nullptr;
        int rc = dev_copy(env, padded.data(), padded.size(), &p);
        if (rc) { cl_destroy(env); return rc; }
        d.table = p;
    }
    {
        std::vector<float> pf((size_t)CL_NPARAM * B);
        for (size_t i = 0; i < pf.size(); ++i) pf[i] = (float)desc->params[i];
        float* p = nullptr; double* q = nullptr; int32_t* ip = nullptr; int4* ds = nullptr;
        int rc = dev_copy(env, pf.data(), pf.size(), &p);
        if (!rc) rc = dev_copy(env, desc->params, (size_t)CL_NPARAM * B, &q);
        if (!rc) rc = dev_copy(env, desc->iparams, (size_t)CL_NIPARAM * B, &ip);
        if (!rc) rc = dev_copy(env, reinterpret_cast<const int4*>(desc->obs_desc), (size_t)d.L, &ds);
        std::vector<int32_t> tcol((size_t)d.L);
        for (int k = 0; k < d.L; ++k) {
            const int32_t* e4 = desc->obs_desc + 4 * (size_t)k;
            if (e4[3] < 0 || e4[3] >= B) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: observation descriptor names a building outside the district"); }
            tcol[k] = e4[0] == CL_OBS_TS ? e4[1] : (e4[0] == CL_OBS_DYN ? -1 : -2 - e4[3]);
            if (e4[0] == CL_OBS_TS && (e4[1] < 0 || e4[1] >= d.W)) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: observation column outside the table"); }
            if (e4[0] == CL_OBS_DYN && (e4[1] < 0 || e4[1] >= CL_NDYN)) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: bad DYN slot"); }
            if (e4[0] == CL_OBS_STATE) {
                const int ncc = desc->ev ? desc->ev->n_constrained : 0;
                if (e4[1] < 0 || e4[1] >= ncc * CL_CC_SLOTS) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: CL_OBS_STATE slot outside the charging-constraint state"); }
                if (!desc->stale_observations) { cl_destroy(env); return fail(CL_ERR_UNSUPPORTED, "cl_create: charging-constraint observations need stale_observations = 1"); }
                d.obs_state = 1;
            } else if (e4[0] != CL_OBS_TS && e4[0] != CL_OBS_DYN && e4[0] != CL_OBS_OUTAGE) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: unknown observation kind"); }
        }
        int32_t* tc = nullptr;
        if (!rc) rc = dev_copy(env, tcol.data(), tcol.size(), &tc);
        if (rc) { cl_destroy(env); return rc; }
        d.pf = p; d.pd = q; d.ip = ip; d.desc = ds; d.tcol = tc;
        std::vector<int32_t> oc;
        for (int k = 0; k < d.L; ++k) if (desc->obs_desc[4 * (size_t)k] == CL_OBS_OUTAGE) oc.push_back(k);
        int32_t* ocd = nullptr;
        rc = dev_copy(env, oc.data(), oc.size(), &ocd);
        if (rc) { cl_destroy(env); return rc; }
        d.out_cols = ocd; d.n_out_cols = (int)oc.size();
        // action-dependent observation columns grouped by building (fresh-observation table path)
        std::vector<int32_t> doff((size_t)B + 1, 0);
        for (int k = 0; k < d.L; ++k) if (desc->obs_desc[4 * (size_t)k] == CL_OBS_DYN) doff[(size_t)desc->obs_desc[4 * (size_t)k + 3] + 1]++;
        for (int b = 0; b < B; ++b) doff[(size_t)b + 1] += doff[(size_t)b];
        std::vector<int2> dcols((size_t)doff[(size_t)B]);
        {
            std::vector<int32_t> fill(doff.begin(), doff.end() - 1);
            for (int k = 0; k < d.L; ++k) {
                const int32_t* e4 = desc->obs_desc + 4 * (size_t)k;
                if (e4[0] == CL_OBS_DYN) dcols[(size_t)fill[(size_t)e4[3]]++] = make_int2(k, e4[1]);
            }
        }
        int2* dcd = nullptr; int32_t* dod = nullptr;
        rc = dev_copy(env, dcols.data(), dcols.size(), &dcd);
        if (!rc) rc = dev_copy(env, doff.data(), doff.size(), &dod);
        if (rc) { cl_destroy(env); return rc; }
        d.dyn_cols = dcd; d.dyn_off = dod;
    }
    {
        void* p = nullptr;
        env->st_floats = (size_t)ST_N * d.U;
        if (cudaMalloc(&p, env->st_floats * sizeof(float)) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: state allocation failed"); }
        env->allocs.push_back(p); d.st = static_cast<float*>(p);
        cudaMemset(p, 0, env->st_floats * sizeof(float));
        env->dst_doubles = env->precision == CL_PRECISION_FP64 ? (size_t)2 * d.U : 0;
        if (env->dst_doubles) {
            if (cudaMalloc(&p, env->dst_doubles * sizeof(double)) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: state allocation failed"); }
            env->allocs.push_back(p); d.dst = static_cast<double*>(p);
            cudaMemset(p, 0, env->dst_doubles * sizeof(double));
        }
        env->lst_floats = any_dyn ? (size_t)kLstmStateFloats * d.U : 0;
        if (env->lst_floats) {
            if (cudaMalloc(&p, env->lst_floats * sizeof(float)) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: LSTM state allocation failed"); }
            env->allocs.push_back(p); d.lst = static_cast<float*>(p);
            cudaMemset(p, 0, env->lst_floats * sizeof(float));
        }
        if (cudaMalloc(&p, (size_t)d.E * sizeof(int32_t)) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: start allocation failed"); }
        env->allocs.push_back(p); d.start = static_cast<int32_t*>(p);
        if (cudaMalloc(&p, sizeof(int32_t)) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: time-step allocation failed"); }
        env->allocs.push_back(p); env->t_dev = static_cast<int32_t*>(p); d.t_dev = env->t_dev;
        cudaMemset(p, 0, sizeof(int32_t));
    }
    if (desc->ev != nullptr && (desc->ev->n_chargers > 0 || desc->ev->n_machines > 0)) {
        // electric vehicles / chargers / washing machines (cl_ev_desc)
        const cl_ev_desc& ev = *desc->ev;
        if (ev.n_ev < 0 || ev.n_chargers < 0 || ev.n_machines < 0) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: negative vehicle / charger / machine count"); }
        if (any_dyn || !d.stale) { cl_destroy(env); return fail(CL_ERR_UNSUPPORTED, "cl_create: districts with vehicles run with reference-parity (stale) observations and without LSTM dynamics"); }
        if (B > 480) { cl_destroy(env); return fail(CL_ERR_UNSUPPORTED, "cl_create: districts with vehicles are not building-tiled (at most 480 buildings)"); }
        if (ev.n_chargers > 0 && (ev.n_ev < 1 || !ev.ev_params || !ev.ev_iparams || !ev.ev_cols || !ev.ev_drift || !ev.ch_building || !ev.ch_action || !ev.ch_cols || !ev.ch_params))
            { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: incomplete cl_ev_desc"); }
        if (ev.n_machines > 0 && (!ev.wm_building || !ev.wm_action || !ev.wm_cols)) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: incomplete cl_ev_desc (washing machines)"); }
        std::vector<int32_t> ch_off((size_t)B + 1, 0), wm_off((size_t)B + 1, 0);
        for (int k = 0; k < ev.n_chargers; ++k) {
            const int b = ev.ch_building[k];
            if (b < 0 || b >= B || (k > 0 && b < ev.ch_building[k - 1])) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: chargers must be listed in building order"); }
            ch_off[(size_t)b + 1]++;
            for (int j = 0; j < 4; ++j) if (ev.ch_cols[k * 4 + j] < 0 || ev.ch_cols[k * 4 + j] >= d.W) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: charger column outside the table"); }
            if (ev.ch_action[k] >= d.A) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: charger action slot outside the action vector"); }
        }
        for (int k = 0; k < ev.n_machines; ++k) {
            const int b = ev.wm_building[k];
            if (b < 0 || b >= B || (k > 0 && b < ev.wm_building[k - 1])) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: washing machines must be listed in building order"); }
            wm_off[(size_t)b + 1]++;
            for (int j = 0; j < 4; ++j) if (ev.wm_cols[k * 4 + j] < 0 || ev.wm_cols[k * 4 + j] >= d.W) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: washing-machine column outside the table"); }
            int max_len = 1;                       // the partial-sum columns follow the load column
            for (int r = 0; r < d.n_rows; ++r) max_len = std::max(max_len, (int)desc->table[(size_t)r * d.W + ev.wm_cols[k * 4 + 3]]);
            if (ev.wm_cols[k * 4 + 2] + max_len - 1 >= d.W) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: washing-machine partial-sum columns outside the table"); }
            if (ev.wm_action[k] >= d.A) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: washing-machine action slot outside the action vector"); }
        }
        for (int b = 0; b < B; ++b) {
            if (ch_off[(size_t)b + 1] > CL_MAX_CHARGERS_PER_BUILDING) { cl_destroy(env); return fail(CL_ERR_UNSUPPORTED, "cl_create: more than CL_MAX_CHARGERS_PER_BUILDING chargers on one building"); }
            ch_off[(size_t)b + 1] += ch_off[(size_t)b]; wm_off[(size_t)b + 1] += wm_off[(size_t)b];
        }
        for (int v = 0; v < ev.n_ev; ++v) {
            if (ev.ev_params[(size_t)v * CL_NPARAM + CL_P_TIME_STEP_RATIO] != 1.0) { cl_destroy(env); return fail(CL_ERR_UNSUPPORTED, "cl_create: vehicles need time_step_ratio == 1"); }
            for (int j = 0; j < 4; ++j) if (j != 2 && (ev.ev_cols[v * 4 + j] < 0 || ev.ev_cols[v * 4 + j] >= d.W)) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: vehicle column outside the table"); }
        }
        double *evp = nullptr, *drift = nullptr, *chp = nullptr;
        int32_t *evip = nullptr, *evc = nullptr, *cho = nullptr, *cha = nullptr, *chc = nullptr, *wmo = nullptr, *wma = nullptr, *wmc = nullptr;
        const int32_t zero4[4] = {0, 0, 0, 0};
        int rc = dev_copy(env, ev.ev_params, (size_t)ev.n_ev * CL_NPARAM, &evp);
        if (!rc) rc = dev_copy(env, ev.ev_iparams, (size_t)ev.n_ev * 2, &evip);
        if (!rc) rc = dev_copy(env, ev.ev_cols, (size_t)ev.n_ev * 4, &evc);
        if (!rc) rc = dev_copy(env, ev.ev_drift, (size_t)d.n_rows * ev.n_ev, &drift);
        if (!rc) rc = dev_copy(env, ch_off.data(), ch_off.size(), &cho);
        if (!rc) rc = dev_copy(env, ev.n_chargers ? ev.ch_action : zero4, (size_t)ev.n_chargers, &cha);
        if (!rc) rc = dev_copy(env, ev.n_chargers ? ev.ch_cols : zero4, (size_t)ev.n_chargers * 4, &chc);
        if (!rc) rc = dev_copy(env, ev.ch_params, (size_t)ev.n_chargers * CL_NCHP, &chp);
        if (!rc) rc = dev_copy(env, wm_off.data(), wm_off.size(), &wmo);
        if (!rc) rc = dev_copy(env, ev.n_machines ? ev.wm_action : zero4, (size_t)ev.n_machines, &wma);
        if (!rc) rc = dev_copy(env, ev.n_machines ? ev.wm_cols : zero4, (size_t)ev.n_machines * 4, &wmc);
        if (rc) { cl_destroy(env); return rc; }
        d.ev_n = ev.n_ev; d.ch_n = ev.n_chargers; d.wm_n = ev.n_machines;
        d.ev_pd = evp; d.ev_ip = evip; d.ev_cols = evc; d.ev_drift = drift; d.ch_off = cho; d.ch_action = cha; d.ch_cols = chc; d.ch_pd = chp;
        d.wm_off = wmo; d.wm_action = wma; d.wm_cols = wmc;
        // charging constraints
        const int ncc = ev.n_constrained > 0 ? ev.n_constrained : 0;
        if (ncc > 0) {
            if (!ev.cc_building || !ev.cc_limits || !ev.cc_members || !ev.cc_flags) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: null charging-constraint arrays"); }
            std::vector<int32_t> idx((size_t)B, -1);
            for (int k = 0; k < ncc; ++k) {
                const int b = ev.cc_building[k];
                if (b < 0 || b >= B || (k > 0 && b <= ev.cc_building[k - 1])) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: constrained buildings must be listed in ascending order"); }
                idx[(size_t)b] = k;
                const int nch = ch_off[(size_t)b + 1] - ch_off[(size_t)b];
                for (int j = 0; j < CL_MAX_PHASES * CL_MAX_CHARGERS_PER_BUILDING; ++j) {
                    const int m = ev.cc_members[(size_t)k * CL_MAX_PHASES * CL_MAX_CHARGERS_PER_BUILDING + j];
                    if (m >= nch || m < -1) { cl_destroy(env); return fail(CL_ERR_INVALID, "cl_create: charging-phase member outside the building's chargers"); }
                }
            }
            int32_t *ci = nullptr, *cm = nullptr, *cf = nullptr; double* cl_ = nullptr;
            int rc2 = dev_copy(env, idx.data(), idx.size(), &ci);
            if (!rc2) rc2 = dev_copy(env, ev.cc_limits, (size_t)ncc * (1 + CL_MAX_PHASES), &cl_);
            if (!rc2) rc2 = dev_copy(env, ev.cc_members, (size_t)ncc * CL_MAX_PHASES * CL_MAX_CHARGERS_PER_BUILDING, &cm);
            if (!rc2) rc2 = dev_copy(env, ev.cc_flags, (size_t)ncc, &cf);
            if (rc2) { cl_destroy(env); return rc2; }
            d.cc_n = ncc; d.cc_index = ci; d.cc_limits = cl_; d.cc_members = cm; d.cc_flags = cf;
        }
        void* q = nullptr;
        const size_t ev_floats = (size_t)2 * d.E * std::max(ev.n_ev, 1);
        env->ev_sf_floats = ev_floats + (size_t)d.E * ncc * CL_CC_SLOTS; env->ev_sd_doubles = ev_floats;
        env->ev_flag_bytes = (size_t)d.E * std::max(ev.n_ev, 1); env->wm_flag_bytes = (size_t)d.E * std::max(ev.n_machines, 1);
        if (cudaMalloc(&q, env->ev_sf_floats * sizeof(float)) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: vehicle state allocation failed"); }
        env->allocs.push_back(q); d.ev_sf = static_cast<float*>(q); cudaMemset(q, 0, env->ev_sf_floats * sizeof(float));
        if (ncc > 0) {        // headroom = the caps, violation 0 until the first applied actions (`_set_default_charging_headroom`); never reset
            d.cc_state = d.ev_sf + ev_floats;
            std::vector<float> init((size_t)d.E * ncc * CL_CC_SLOTS, 0.f);
            for (int e = 0; e < d.E; ++e)
                for (int k = 0; k < ncc; ++k)
                    for (int j = 0; j <= CL_MAX_PHASES; ++j) {
                        const double v = ev.cc_limits[(size_t)k * (1 + CL_MAX_PHASES) + j];
                        init[((size_t)e * ncc + k) * CL_CC_SLOTS + j] = v == v ? (float)v : 0.f;
                    }
            if (cudaMemcpy(d.cc_state, init.data(), init.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: charging-constraint state upload failed"); }
        }
        if (cudaMalloc(&q, env->ev_sd_doubles * sizeof(double)) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: vehicle state allocation failed"); }
        env->allocs.push_back(q); d.ev_sd = static_cast<double*>(q); cudaMemset(q, 0, env->ev_sd_doubles * sizeof(double));
        if (cudaMalloc(&q, env->ev_flag_bytes) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: vehicle state allocation failed"); }
        env->allocs.push_back(q); d.ev_flag = static_cast<uint8_t*>(q); cudaMemset(q, 0, env->ev_flag_bytes);
        if (cudaMalloc(&q, env->wm_flag_bytes) != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, "cl_create: washing-machine state allocation failed"); }
        env->allocs.push_back(q); d.wm_flag = static_cast<uint8_t*>(q); cudaMemset(q, 0, env->wm_flag_bytes);
    }
    {
    }
    // launch geometry.  A block = whole envs x all B buildings (physics threads, rounded up to warps) + one helper warp.
    // Every block must be RESIDENT for the whole launch to run in one wave (the kernel is register-heavy: ~100-128 registers
    // per thread), so the block size is chosen per (E, B, device): among a few candidate sizes pick the one that minimises
    // waves x resident threads per SM.  On B200 at 17 x 4096 this selects one 512-thread block per SM (147 blocks on 148 SMs;
    // profiles/README.md has the sweep).  CL_B200_BLOCK_THREADS overrides for experiments.
    int n_sm = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm < 1) n_sm = 148;
    }
    env->n_sm = n_sm;
    int regs = 128;
    {
        cudaFuncAttributes fa;
        const void* fn = env->precision == CL_PRECISION_FP64
            ? (any_dyn ? (const void*)advance_kernel<double, true, true, kDynMaxT> : (any_thermal ? (const void*)advance_kernel<double, true, false, 512> : (const void*)advance_kernel<double, false, false, 512>))
            : (any_dyn ? (const void*)advance_kernel<float, true, true, kDynMaxT> : (any_thermal ? (const void*)advance_kernel<float, true, false, 512> : (const void*)advance_kernel<float, false, false, 512>));
        if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess && fa.numRegs > 0) regs = fa.numRegs;
        cudaGetLastError();
    }
    // threads per env: one per building
    const int BT = B;
    int target = 0;
    if (const char* ev = std::getenv("CL_B200_BLOCK_THREADS")) { const int v = std::atoi(ev); if (v >= 32 && v <= 992) target = v; }
    if (target == 0) {
        const int cand[] = {96, 128, 224, 352, 480};
        double best = 1e30;
        for (int ci = 0; ci < 5; ++ci) {
            int epb_c = cand[ci] / BT;
            if (epb_c < 1) epb_c = 1;
            if (d.lstm_smem == 2 && epb_c >= 32) epb_c &= ~31;       // tensor-core LSTM cell: whole warps on one building
            if (epb_c > d.E) epb_c = d.E;
            const int thr = ((epb_c * BT + 31) / 32) * 32 + 32;
            if (thr > (any_dyn ? kDynMaxT : 512) && ci > 0) continue;
            const int regs_alloc = ((regs + 7) / 8) * 8;
            int bps = 65536 / (regs_alloc * thr);
            if (bps > 2048 / thr) bps = 2048 / thr;
            if (bps < 1) bps = 1;
            if (d.lstm_smem == 2) {
                // the operand fragments + packed weights + projections of the tensor-core LSTM cell are per block: shared memory, not
                // registers, bounds the resident blocks (measured: 2 waves of 224-thread blocks at 3 x 16 384 took 0.31 ms/step)
                const size_t fixed = (size_t)B * (kLstmStride + kLstmFragFloats + kLstmPreRing * 64) * sizeof(float) + 16 * 1024;
                bps = std::max(1, std::min(bps, (int)((227 * 1024) / fixed)));
            }
            const long blocks = (d.E + epb_c - 1) / epb_c;
            const long waves = (blocks + (long)n_sm * bps - 1) / ((long)n_sm * bps);
            const long resident = blocks < (long)n_sm * bps ? (blocks + n_sm - 1) / n_sm : bps;   // blocks per SM actually resident
            // (LSTM blocks are latency bound: a wave costs about the same whatever its width - fewest waves first, then the widest block)
            const double cost = d.lstm_smem == 2 ? (double)waves * 1e6 - (double)cand[ci] : (double)waves * (double)resident * thr;
            // (ties go to the larger block: per-block staging - curves, LSTM weights - is amortised over more units; measured on the
            // LSTM district: 1.18 ms/step with 512-thread blocks against 1.35 with two 256-thread blocks per SM)
            if (cost < best - 1e-9 || (cost < best + 1e-9 && cand[ci] > target)) { best = cost; target = cand[ci]; }
        }
        if (target == 0) target = 128;
        // Many more units than one wave of 512-thread blocks can hold (e.g. 17 x 32768): the 1024-thread instantiation (64
        // registers, a few spills) keeps twice the warps resident per SM and wins by 10 % (fp64 flow) / 16 % (fp32) there
        // (profiles/README.md "Large env counts"); thermal / LSTM instantiations spill too much at 64 registers.
        if (!any_thermal && !any_dyn && B <= 480 && d.stale) {      // (per-env row images of fresh observations need the 512-thread geometry)
            const int epb_t = std::max(1, std::min(target / B, d.E));
            const long blocks_t = (d.E + epb_t - 1) / epb_t;
            const int epb_big = std::max(1, 960 / B);
            const long blocks_big = (d.E + epb_big - 1) / epb_big;
            if (blocks_t > 2L * n_sm && blocks_big >= n_sm) target = 960;
        }
    }
    int epb = target / BT;
    if (epb < 1) epb = 1;
    if (d.lstm_smem == 2 && epb >= 32) epb &= ~31;
    if (epb > d.E) epb = d.E;
    d.envs_per_block = epb;
    d.tiles = 1; d.tile_b = B; d.Lt = d.L; d.tile_k = nullptr;
    const bool tab_fits = obs_table_fits(d);
    d.tab_layout = tab_fits ? 1 : 0;
    env->threads = ((epb * BT + 31) / 32) * 32;
    env->blocks = (d.E + epb - 1) / epb;
    // observations with action-dependent columns: per-env row images (2 x envs_per_block x L floats of shared memory) when they fit
    // and are large enough to host the general writer's dynbuf, which aliases them (smem_layout)
    auto fresh_fits = [&](const Dev& q, int nthreads) -> bool {
        if (q.stale || !q.tab_layout) return false;
        Dev t = q; t.fresh_slots = 1;
        const int lp = (t.Lt + 3) & ~3;
        if ((size_t)2 * t.envs_per_block * lp < (size_t)nthreads * CL_NDYN) return false;
        return smem_bytes(t, nthreads, false, env->precision == CL_PRECISION_FP64 ? 8 : 4) <= 200 * 1024;
    };
    if (!d.stale) { d.fresh_slots = fresh_fits(d, env->threads + 32) ? 1 : 0; d.tab_layout = d.fresh_slots; }
    // Wide districts: one env per thread-block CLUSTER, the buildings split into `tiles` tiles of `tile_b` (one CTA each); the
    // district sums travel through distributed shared memory.  Chosen when a whole env does not fit one 512-thread block
    // (CL_B200_TILES forces a tile count for experiments).
    int want_tiles = 0;
    if (const char* ev = std::getenv("CL_B200_TILES")) { const int v = std::atoi(ev); if (v >= 2 && v <= 8) want_tiles = v; }
    if ((B > 480 || want_tiles) && !any_dyn) {
        int regs_w = regs;
        {
            cudaFuncAttributes fa;
            const void* fn = env->precision == CL_PRECISION_FP64
                ? (any_thermal ? (const void*)advance_kernel<double, true, false, 512, true> : (const void*)advance_kernel<double, false, false, 512, true>)
                : (any_thermal ? (const void*)advance_kernel<float, true, false, 512, true> : (const void*)advance_kernel<float, false, false, 512, true>);
            if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess && fa.numRegs > 0) regs_w = fa.numRegs;
            cudaGetLastError();
        }
        // observation columns of a tile must be one contiguous range: rows are ordered by building (schema.observation_layout)
        auto tile_ranges = [&](int nt_, int tb, std::vector<int32_t>& tk) -> bool {
            tk.assign(nt_ + 1, d.L);
            int prev = 0;
            tk[0] = 0;
            for (int k = 0; k < d.L; ++k) {
                const int bb = desc->obs_desc[4 * (size_t)k + 3];
                if (bb < prev) return false;
                for (int r = prev / tb + 1; r <= bb / tb && r <= nt_; ++r) tk[r] = k;
                prev = bb;
            }
            for (int r = 1; r <= nt_; ++r) if (tk[r] < tk[r - 1]) tk[r] = tk[r - 1];
            tk[nt_] = d.L;
            return true;
        };
        int best_nt = 0; double best_score = -1.0; std::vector<int32_t> best_tk;
        for (int nt_ = 2; nt_ <= 8; ++nt_) {
            if (want_tiles && nt_ != want_tiles) continue;
            const int tb = (B + nt_ - 1) / nt_;
            if (tb > 480 || (nt_ - 1) * tb >= B) continue;              // every tile must own at least one building
            std::vector<int32_t> tk;
            if (!tile_ranges(nt_, tb, tk)) continue;
            int lt = 0;
            for (int r = 0; r < nt_; ++r) lt = std::max(lt, tk[r + 1] - tk[r]);
            const int thr = ((tb + 31) / 32) * 32 + 32;
            bool aligned = true;
            for (int r = 0; r <= nt_; ++r) aligned = aligned && (tk[r] & 3) == 0;
            Dev probe = d; probe.tiles = nt_; probe.tile_b = tb; probe.Lt = lt; probe.envs_per_block = 1; probe.fresh_slots = 0;
            probe.tab_layout = (tab_fits && aligned) ? 1 : 0;
            // (the general writer's dynbuf only exists with action-dependent observation columns)
            const size_t sm = smem_bytes(probe, thr, !d.stale, env->precision == CL_PRECISION_FP64 ? 8 : 4);
            if (sm > 200 * 1024) continue;
            if (smem_bytes(probe, thr - 32, true, env->precision == CL_PRECISION_FP64 ? 8 : 4) > 227 * 1024) continue;   // the reset kernel's staging
            const int regs_alloc = ((regs_w + 7) / 8) * 8;
            int bps = 65536 / (regs_alloc * thr);
            bps = std::min(bps, 2048 / thr);
            bps = std::min(bps, (int)((227 * 1024) / (sm + 1024)));
            if (bps < 1) continue;
            const double score = (double)bps * tb - 1e-3 * nt_;           // resident physics threads per SM; fewer tiles on ties
            if (score > best_score) { best_score = score; best_nt = nt_; best_tk = tk; }
        }
        if (best_nt == 0) {
            if (B > 992 || want_tiles) { cl_destroy(env); return fail(CL_ERR_UNSUPPORTED, "cl_create: no building tiling fits this district (observation rows must be ordered by building; <= 8 tiles of <= 480 buildings)"); }
        } else {
            const int tb = (B + best_nt - 1) / best_nt;
            int lt = 0;
            for (int r = 0; r < best_nt; ++r) lt = std::max(lt, best_tk[r + 1] - best_tk[r]);
            int32_t* tkd = nullptr;
            int rc = dev_copy(env, best_tk.data(), best_tk.size(), &tkd);
            if (rc) { cl_destroy(env); return rc; }
            bool aligned = true;
            for (int r = 0; r <= best_nt; ++r) aligned = aligned && (best_tk[r] & 3) == 0;
            d.tab_layout = (tab_fits && aligned) ? 1 : 0;
            d.tiles = best_nt; d.tile_b = tb; d.Lt = lt; d.tile_k = tkd; d.envs_per_block = 1;
            env->wide = true;
            env->threads = ((tb + 31) / 32) * 32;
            env->blocks = d.E * best_nt;
            d.fresh_slots = 0;
            if (!d.stale) { d.fresh_slots = fresh_fits(d, env->threads + 32) ? 1 : 0; d.tab_layout = d.fresh_slots; }
        }
    }
    if (!env->wide && B > 992) { cl_destroy(env); return fail(CL_ERR_UNSUPPORTED, "cl_create: districts wider than 992 buildings with LSTM dynamics are not supported"); }
    // opt in to large dynamic shared memory once (both kernels, all instantiations)
    const size_t smem = smem_bytes(d, env->threads + 32, !d.stale, env->precision == CL_PRECISION_FP64 ? 8 : 4);   // (dynbuf: only with action-dependent observation columns)
    if (smem > 200 * 1024) { cl_destroy(env); return fail(CL_ERR_UNSUPPORTED, "cl_create: district too wide for the shared-memory staging"); }
    // (the reset kernel stages every unit's t = 0 values through the general writer's dynbuf)
    const size_t smem_reset = smem_bytes(d, env->threads, true, env->precision == CL_PRECISION_FP64 ? 8 : 4);
    if (smem_reset > 227 * 1024) { cl_destroy(env); return fail(CL_ERR_UNSUPPORTED, "cl_create: district too wide for the shared-memory staging"); }
    ensure_smem_optin(std::max<size_t>(std::max(smem, smem_reset), 116 * 1024));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cl_destroy(env); return fail(CL_ERR_CUDA, std::string("cl_create: ") + cudaGetErrorString(e)); }
    { const int rc = build_obs_table(env); if (rc) { cl_destroy(env); return rc; } }
    *out = env;
    return CL_OK;
}

extern "C" int cl_destroy(cl_env* env) {
    if (!env) return CL_OK;
    for (void* p : env->allocs) cudaFree(p);
    if (env->outage_dev) cudaFree(env->outage_dev);
    if (env->dpart) cudaFree(env->dpart);
    if (env->kpi_unit) cudaFree(env->kpi_unit);
    if (env->kpi_env) cudaFree(env->kpi_env);
    for (int i = 0; i < 64; ++i) if (g_lstm_const_owner[i] == env) g_lstm_const_owner[i] = nullptr;
    for (void* q : env->x_opened) cudaIpcCloseMemHandle(q);
    if (env->x_buf) cudaFree(env->x_buf);
    if (env->x_peers_dev) cudaFree(env->x_peers_dev);
    if (env->x_err_dev) cudaFree(env->x_err_dev);
    if (env->host_flag) cudaFreeHost(const_cast<int32_t*>(env->host_flag));
    delete env;
    return CL_OK;
}

extern "C" int cl_set_outage(cl_env* env, const float* signals, int32_t episode_time_steps, cl_stream stream) {
    if (!env) return fail(CL_ERR_INVALID, "cl_set_outage: null env");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!signals) { env->d.has_outage = 0; env->d.outage = nullptr; return CL_OK; }
    if (episode_time_steps < 1) return fail(CL_ERR_INVALID, "cl_set_outage: bad episode_time_steps");
    const size_t n = (size_t)env->d.B * episode_time_steps;
    if (!env->outage_dev || env->outage_T != episode_time_steps) {
        if (env->outage_dev) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(env->outage_dev); env->outage_dev = nullptr; }
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->outage_dev), n * sizeof(float)));
        env->outage_T = episode_time_steps;
    }
    CUDA_TRY(cudaMemcpyAsync(env->outage_dev, signals, n * sizeof(float), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // `signals` is a host buffer the caller may free right away
    env->d.outage = env->outage_dev;
    env->d.has_outage = 1;
    return CL_OK;
}

// blocks of up to 512 threads get the 128-register budget; only very wide districts (B > 512) use 1024-thread blocks
template <typename R, bool TH>
static void launch_reset(cl_env* env, float* obs, cudaStream_t st) {
    const size_t smem = smem_bytes(env->d, env->threads, true, (int)sizeof(R));
    if (env->threads <= 512) reset_kernel<R, TH, 512><<<env->blocks, env->threads, smem, st>>>(env->d, obs);
    else reset_kernel<R, TH, 1024><<<env->blocks, env->threads, smem, st>>>(env->d, obs);
}
template <typename R, bool TH, bool DY>
static void launch_advance(cl_env* env, int t0, int K, const float* actions, float* obs, float* reward, float* district, float* trace,
                           bool coupled, cudaStream_t st) {
    const bool want_dyn = !env->d.stale && obs != nullptr;
    const int nthreads = env->threads + 32;          // + the helper warp
    size_t smem = smem_bytes(env->d, nthreads, want_dyn, (int)sizeof(R));
    // a launch of at most one block per SM must SPREAD over the SMs: when two of its blocks would fit one SM (small blocks, the float
    // instantiations) the block scheduler may pair them up and leave SMs idle (measured: 17 x 4096 fp32 3.6 -> 5.4 us / step) - ask
    // for more than half an SM's shared memory so that a block owns its SM
    if (env->blocks <= env->n_sm && smem < 116 * 1024) smem = 116 * 1024;
    if (env->wide) {
        if constexpr (!DY) {
            Dev dd = env->d;
            dd.coupled = coupled ? 1 : 0;
            if (coupled) {
                // tiles exchange sums inside the step: one thread-block cluster per env, one CTA per building tile
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)env->blocks); cfg.blockDim = dim3((unsigned)nthreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = (unsigned)env->d.tiles; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, advance_kernel<R, TH, false, 512, true>, dd, t0, K, actions, obs, reward, district, trace);
            } else {
                advance_kernel<R, TH, false, 512, true><<<env->blocks, nthreads, smem, st>>>(dd, t0, K, actions, obs, reward, district, trace);
            }
        }
        return;
    }
    if constexpr (!DY) {
        if (env->d.ch_n + env->d.wm_n > 0) {          // districts with electric vehicles / washing machines (cl_ev_desc)
            if (nthreads <= 512) advance_kernel<R, TH, false, 512, false, false, true><<<env->blocks, nthreads, smem, st>>>(env->d, t0, K, actions, obs, reward, district, trace);
            else advance_kernel<R, TH, false, 1024, false, false, true><<<env->blocks, nthreads, smem, st>>>(env->d, t0, K, actions, obs, reward, district, trace);
            return;
        }
        {
            // nothing in the step reads the district sum and every env is on the same row: no block barrier per step (DEC).  OPT-IN
            // (CL_B200_DECOUPLE=1 at cl_create): measured on B200 it is 1.5 - 3.5 % SLOWER than the plain block barrier
            const Dev& q = env->d;
            const bool dsum_in_step = q.reward_id == CL_REWARD_MARL || q.reward_id == CL_REWARD_ELECTRIC_VEHICLES || q.central;
            if (env->decouple && !env->kpi_fused && !dsum_in_step && q.stale && q.uniform_start && q.x_n <= 1 && !q.obs_state) {
                if (nthreads <= 512) advance_kernel<R, TH, false, 512, false, false, false, true><<<env->blocks, nthreads, smem, st>>>(env->d, t0, K, actions, obs, reward, district, trace);
                else advance_kernel<R, TH, false, 1024, false, false, false, true><<<env->blocks, nthreads, smem, st>>>(env->d, t0, K, actions, obs, reward, district, trace);
                return;
            }
        }
        if (env->kpi_fused) {                         // online KPI accumulators inside the step (cl_kpi_enable)
            if (nthreads <= 512) advance_kernel<R, TH, false, 512, false, true><<<env->blocks, nthreads, smem, st>>>(env->d, t0, K, actions, obs, reward, district, trace);
            else advance_kernel<R, TH, false, 1024, false, true><<<env->blocks, nthreads, smem, st>>>(env->d, t0, K, actions, obs, reward, district, trace);
            return;
        }
    }
    if (nthreads <= (DY ? kDynMaxT : 512)) advance_kernel<R, TH, DY, (DY ? kDynMaxT : 512)><<<env->blocks, nthreads, smem, st>>>(env->d, t0, K, actions, obs, reward, district, trace);
    else advance_kernel<R, TH, DY, 1024><<<env->blocks, nthreads, smem, st>>>(env->d, t0, K, actions, obs, reward, district, trace);
}
// The constant bank holds ONE district's LSTM weights per device: the handle that launches next claims it (device-wide synchronisation
// + 51 KB copy when the owner changes; never inside a stream capture - the launch then uses the shared-memory path).
static void claim_lstm_constants(cl_env* env, cudaStream_t st) {
    env->d.lstm_const = 0;
    if (env->lstm_packed.empty()) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    if (g_lstm_const_owner[dev] != env) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
        if (cudaDeviceSynchronize() != cudaSuccess) return;
        if (cudaMemcpyToSymbol(c_lstm_w, env->lstm_packed.data(), env->lstm_packed.size() * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return; }
        g_lstm_const_owner[dev] = env;
    }
    env->d.lstm_const = 1;
}

static void dispatch_one(cl_env* env, int t0, int K, const float* actions, float* obs, float* reward, float* district, float* trace,
                         bool coupled, cudaStream_t st) {
    if (env->d.x_n > 1) { env->d.x_epoch = env->x_epoch; env->x_epoch += (unsigned)K; }   // every rank runs the same launch sequence
    if (env->dynamics) claim_lstm_constants(env, st);
    if (env->precision == CL_PRECISION_FP64) {
        if (env->dynamics) launch_advance<double, true, true>(env, t0, K, actions, obs, reward, district, trace, coupled, st);
        else if (env->thermal) launch_advance<double, true, false>(env, t0, K, actions, obs, reward, district, trace, coupled, st);
        else launch_advance<double, false, false>(env, t0, K, actions, obs, reward, district, trace, coupled, st);
    } else {
        if (env->dynamics) launch_advance<float, true, true>(env, t0, K, actions, obs, reward, district, trace, coupled, st);
        else if (env->thermal) launch_advance<float, true, false>(env, t0, K, actions, obs, reward, district, trace, coupled, st);
        else launch_advance<float, false, false>(env, t0, K, actions, obs, reward, district, trace, coupled, st);
    }
    env->launches++;
}
static int dispatch_advance(cl_env* env, int K, const float* actions, float* obs, float* reward, float* district, float* trace, cudaStream_t st) {
    const Dev& d = env->d;
    // same predicate as the kernel's need_dsum || central_sync
    const bool coupled = env->wide && reward != nullptr && (d.reward_id == CL_REWARD_MARL || (d.central && d.reward_id >= 0));
    if (!env->wide || coupled || district == nullptr) {
        dispatch_one(env, env->t, K, actions, obs, reward, district, trace, coupled, st);
        return CL_OK;
    }
    // wide district with independent tiles: the kernel leaves per-tile partial district sums in a scratch buffer, a second small
    // kernel adds the tiles up.  Long rollouts run in chunks so that the scratch stays small ([chunk][E][tiles][3] floats).
    const int chunk = 64;
    const size_t need = (size_t)std::min(K, chunk) * d.E * d.tiles * 3;
    if (env->dpart_floats < need) {
        if (env->dpart) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(env->dpart); env->dpart = nullptr; env->dpart_floats = 0; }
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->dpart), need * sizeof(float)));
        env->dpart_floats = need;
    }
    const int Rdim = d.central ? 1 : d.B;
    for (int off = 0; off < K; off += chunk) {
        const int kc = std::min(chunk, K - off);
        dispatch_one(env, env->t + off, kc, actions + (size_t)off * d.E * d.A, obs ? obs + (size_t)off * d.E * d.L : nullptr,
                     reward ? reward + (size_t)off * d.E * Rdim : nullptr, env->dpart, trace, false, st);
        const long n3 = (long)kc * d.E * 3;
        district_finish_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, st>>>(env->dpart, district + (size_t)off * d.E * 3, n3, d.tiles);
        env->launches++;
    }
    return CL_OK;
}

// device-resident time step: t <- v, or t += k while the episode lasts (the launch that follows a finished episode is a no-op too)
__global__ void set_time_kernel(int32_t* t, int v) { *t = v; }
__global__ void bump_time_kernel(int32_t* t, int k, int T) { if (*t + k <= T - 1) *t += k; }

__global__ void fill_start_kernel(int32_t* start, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) start[i] = v;
}

// device-time mode: the counter on the device is authoritative (a captured graph may have advanced it any number of times)
static int refresh_time(cl_env* env, cudaStream_t st) {
    if (!env->device_time) return CL_OK;
    int32_t t = 0;
    CUDA_TRY(cudaMemcpyAsync(&t, env->t_dev, sizeof(t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    env->t = t;
    return CL_OK;
}
static int publish_time(cl_env* env, cudaStream_t st) {
    if (!env->device_time) return CL_OK;
    set_time_kernel<<<1, 1, 0, st>>>(env->t_dev, env->t);
    CUDA_TRY(cudaGetLastError());
    env->launches++;
    return CL_OK;
}

extern "C" int cl_reset(cl_env* env, const int32_t* episode_start, int32_t uniform_start, int32_t episode_time_steps, float* obs,
                        cl_stream stream) {
    if (!env) return fail(CL_ERR_INVALID, "cl_reset: null env");
    Dev& d = env->d;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (episode_time_steps < 2) return fail(CL_ERR_INVALID, "cl_reset: episode_time_steps must be >= 2");
    // the kernels index the signal with the episode length as the row stride
    if (d.has_outage && env->outage_T != episode_time_steps) return fail(CL_ERR_STATE, "cl_reset: the outage signal must have exactly episode_time_steps entries per building (cl_set_outage)");
    if (episode_start == nullptr) {
        if (uniform_start < 0 || uniform_start + episode_time_steps > d.n_rows) return fail(CL_ERR_INVALID, "cl_reset: episode window outside the table");
        fill_start_kernel<<<(d.E + 255) / 256, 256, 0, st>>>(const_cast<int32_t*>(d.start), d.E, uniform_start);
        env->launches++;
        d.uniform_start = 1; d.start0 = uniform_start;
    } else {
        if (d.ch_n + d.wm_n > 0) return fail(CL_ERR_UNSUPPORTED, "cl_reset: districts with electric vehicles / washing machines use one episode window for all envs");
        CUDA_TRY(cudaMemcpyAsync(const_cast<int32_t*>(d.start), episode_start, sizeof(int32_t) * d.E, cudaMemcpyDeviceToDevice, st));
        d.uniform_start = 0; d.start0 = 0;
    }
    d.T = episode_time_steps;
    env->T = episode_time_steps;
    if (env->precision == CL_PRECISION_FP64) { if (env->thermal) launch_reset<double, true>(env, obs, st); else launch_reset<double, false>(env, obs, st); }
    else { if (env->thermal) launch_reset<float, true>(env, obs, st); else launch_reset<float, false>(env, obs, st); }
    env->launches++;
    CUDA_TRY(cudaGetLastError());
    if (env->kpi_unit) {
        CUDA_TRY(cudaMemsetAsync(env->kpi_unit, 0, (size_t)d.U * CL_NKPI_UNIT * sizeof(double), st));
        CUDA_TRY(cudaMemsetAsync(env->kpi_env, 0, (size_t)d.E * 2 * CL_NKPI_ENV * sizeof(double), st));
    }
    env->t = 0;
    return publish_time(env, st);
}

extern "C" int cl_step(cl_env* env, const float* actions, float* obs, float* reward, float* district, float* trace, cl_stream stream) {
    if (!env || !actions) return fail(CL_ERR_INVALID, "cl_step: null argument");
    if (env->t < 0) return fail(CL_ERR_STATE, "cl_step: call cl_reset first");
    { const int rc = refresh_time(env, static_cast<cudaStream_t>(stream)); if (rc) return rc; }
    if (env->t >= env->T - 1) return fail(CL_ERR_STATE, "cl_step: episode has ended (terminated); call cl_reset");
    { const int rc = dispatch_advance(env, 1, actions, obs, reward, district, trace, static_cast<cudaStream_t>(stream)); if (rc) return rc; }
    CUDA_TRY(cudaGetLastError());
    env->t += 1;
    return publish_time(env, static_cast<cudaStream_t>(stream));
}

// One step end to end from host buffers in ONE call: H2D of the actions, the step kernel, (optionally) the observation row all envs
// share, ONE D2H of the caller's result range, stream synchronisation.  Every buffer is the caller's (pinned host memory makes both
// copies asynchronous DMA transfers).
// completion signal of cl_step_host: the last (one-thread) kernel of the call stores the call's sequence number into page-locked host
// memory; the host polls it instead of paying cudaStreamSynchronize's wake-up latency
__global__ void signal_kernel(volatile int32_t* flag, int32_t value) {
    __threadfence_system();
    *flag = value;
}

// page-locked host memory the device can address in place (cudaHostAlloc / cudaHostRegister under unified addressing)
static bool device_addressable_host(const void* p, const void** dev_ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return false;
    *dev_ptr = a.devicePointer;
    return true;
}

extern "C" int cl_step_host(cl_env* env, const float* actions_host, float* actions_dev, float* obs_dev, float* reward_dev, float* district_dev,
                            float* row_dev, const void* d2h_src_dev, void* d2h_dst_host, size_t d2h_bytes, int32_t in_place, cl_stream stream) {
    if (!env || !actions_host || !actions_dev) return fail(CL_ERR_INVALID, "cl_step_host: null argument");
    if (d2h_bytes && (!d2h_src_dev || !d2h_dst_host)) return fail(CL_ERR_INVALID, "cl_step_host: null result range");
    if (env->t < 0) return fail(CL_ERR_STATE, "cl_step_host: call cl_reset first");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    { const int rc = refresh_time(env, st); if (rc) return rc; }
    if (env->t >= env->T - 1) return fail(CL_ERR_STATE, "cl_step_host: episode has ended (terminated); call cl_reset");
    if (row_dev && (!env->d.stale || !env->d.uniform_start || env->d.obs_state))
        return fail(CL_ERR_STATE, "cl_step_host: the observation row is only shared by all envs with stale_observations, one episode window and no per-env state columns");
    // in place: the kernel reads the actions from, and writes the results that lie inside the result range to, page-locked host memory
    // directly (loads / posted stores over PCIe inside the step instead of a DMA copy before and after it).  Only when both host
    // ranges are device-addressable and the [E, L] observation slab is not requested (8 MB of stores belong on the copy engine).
    const float* act = actions_dev;
    const void *in_dev = nullptr, *out_dev = nullptr;
    const bool read_in_place = (in_place & 1) && device_addressable_host(actions_host, &in_dev);
    bool direct = (in_place & 2) && obs_dev == nullptr && d2h_bytes != 0 && device_addressable_host(d2h_dst_host, &out_dev);
    if (read_in_place) act = static_cast<const float*>(in_dev);
    if (direct) {
        auto translate = [&](float* p) -> float* {
            const char* lo = static_cast<const char*>(d2h_src_dev);
            const char* q = reinterpret_cast<const char*>(p);
            if (p && d2h_bytes && q >= lo && q < lo + d2h_bytes) return reinterpret_cast<float*>(const_cast<char*>(static_cast<const char*>(out_dev)) + (q - lo));
            return p;
        };
        float* r2 = translate(reward_dev); float* w2 = translate(row_dev);
        // everything inside the range must be produced by this call, or the host range would miss it
        const size_t produced = (r2 != reward_dev ? sizeof(float) * (size_t)env->d.E * (env->d.central ? 1 : env->d.B) : 0) + (w2 != row_dev ? sizeof(float) * (size_t)env->d.L : 0);
        if (produced == d2h_bytes) { reward_dev = r2; row_dev = w2; d2h_bytes = 0; } else direct = false;
    }
    if (!read_in_place) CUDA_TRY(cudaMemcpyAsync(actions_dev, actions_host, sizeof(float) * (size_t)env->d.E * (size_t)std::max(env->d.A, 1), cudaMemcpyHostToDevice, st));
    { const int rc = dispatch_advance(env, 1, act, obs_dev, reward_dev, district_dev, nullptr, st); if (rc) return rc; }
    if (row_dev) {
        const long total = env->d.L;
        obs_rows_kernel<<<(unsigned)std::min<long>((total + 255) / 256, 4096), 256, 0, st>>>(env->d, env->t + 1, 1, row_dev);
        env->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    if (d2h_bytes) CUDA_TRY(cudaMemcpyAsync(d2h_dst_host, d2h_src_dev, d2h_bytes, cudaMemcpyDeviceToHost, st));
    env->t += 1;
    { const int rc = publish_time(env, st); if (rc) return rc; }
    if ((in_place & 4) && env->host_flag == nullptr) {
        void* hf = nullptr;
        if (cudaHostAlloc(&hf, 64, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); hf = nullptr; }
        env->host_flag = static_cast<volatile int32_t*>(hf);
        if (env->host_flag) *env->host_flag = 0;
    }
    if ((in_place & 4) && env->host_flag != nullptr) {
        const int32_t seq = ++env->host_seq;
        signal_kernel<<<1, 1, 0, st>>>(env->host_flag, seq);
        env->launches++;
        CUDA_TRY(cudaGetLastError());
        volatile int32_t* f = env->host_flag;
        for (long spins = 0; *f != seq; ++spins) {
            if ((spins & 0xFFFFF) == 0xFFFFF && cudaStreamQuery(st) != cudaErrorNotReady) break;     // finished, or failed: let the sync below report it
        }
        if (*f != seq) CUDA_TRY(cudaStreamSynchronize(st));
        return CL_OK;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return CL_OK;
}

extern "C" int cl_rollout(cl_env* env, int32_t n_steps, const float* actions, float* obs, float* reward, float* district, cl_stream stream) {
    if (!env || !actions) return fail(CL_ERR_INVALID, "cl_rollout: null argument");
    if (n_steps < 1) return fail(CL_ERR_INVALID, "cl_rollout: n_steps must be >= 1");
    if (env->t < 0) return fail(CL_ERR_STATE, "cl_rollout: call cl_reset first");
    { const int rc = refresh_time(env, static_cast<cudaStream_t>(stream)); if (rc) return rc; }
    if (env->t + n_steps > env->T - 1) return fail(CL_ERR_STATE, "cl_rollout: block runs past the end of the episode");
#ifdef CL_PHASE_TIMING
    { const int rc = dispatch_advance(env, n_steps, actions, obs, reward, nullptr, district, static_cast<cudaStream_t>(stream)); if (rc) return rc; }   // `district` receives the stamps
#else
    { const int rc = dispatch_advance(env, n_steps, actions, obs, reward, district, nullptr, static_cast<cudaStream_t>(stream)); if (rc) return rc; }
#endif
    CUDA_TRY(cudaGetLastError());
    env->t += n_steps;
    return publish_time(env, static_cast<cudaStream_t>(stream));
}

extern "C" int cl_device_time_enable(cl_env* env, cl_stream stream) {
    if (!env) return fail(CL_ERR_INVALID, "cl_device_time_enable: null env");
    if (env->t < 0) return fail(CL_ERR_STATE, "cl_device_time_enable: call cl_reset first");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (env->wide) {
        // independent tiles leave per-tile district partials in a scratch buffer: allocate it now, never inside a capture
        const size_t need = (size_t)64 * env->d.E * env->d.tiles * 3;
        if (env->dpart_floats < need) {
            if (env->dpart) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(env->dpart); env->dpart = nullptr; env->dpart_floats = 0; }
            CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->dpart), need * sizeof(float)));
            env->dpart_floats = need;
        }
    }
    env->device_time = true;
    return publish_time(env, st);
}

extern "C" int cl_advance_device(cl_env* env, int32_t n_steps, const float* actions, float* obs, float* reward, float* district, cl_stream stream) {
    if (!env || !actions) return fail(CL_ERR_INVALID, "cl_advance_device: null argument");
    if (n_steps < 1) return fail(CL_ERR_INVALID, "cl_advance_device: n_steps must be >= 1");
    if (!env->device_time) return fail(CL_ERR_STATE, "cl_advance_device: call cl_device_time_enable first (outside any stream capture)");
    if (env->d.x_n > 1) return fail(CL_ERR_UNSUPPORTED, "cl_advance_device: building-sharded districts advance through cl_step / cl_rollout (the exchange epoch is a launch argument)");
    if (env->wide && n_steps > 64 && district != nullptr) return fail(CL_ERR_UNSUPPORTED, "cl_advance_device: at most 64 steps per launch for building-tiled districts with district sums");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int t_host = env->t;
    env->t = -1;                                     // launches read the device counter
    const int rc = dispatch_advance(env, n_steps, actions, obs, reward, district, nullptr, st);
    env->t = t_host;
    if (rc) return rc;
    bump_time_kernel<<<1, 1, 0, st>>>(env->t_dev, n_steps, env->T);
    env->launches++;
    CUDA_TRY(cudaGetLastError());
    return CL_OK;
}

// ---- building-sharded districts: all-gather of partial district sums through peer memory ----------------------------------------
static size_t exchange_bytes(const cl_env* env, int n) { return (size_t)2 * n * env->d.E * 3 * sizeof(uint2); }

extern "C" int cl_exchange_create(cl_env* env, int32_t n_ranks, int32_t rank, void* ipc_handle_out, void** buffer_out) {
    if (!env) return fail(CL_ERR_INVALID, "cl_exchange_create: null env");
    if (n_ranks < 2 || n_ranks > 64 || rank < 0 || rank >= n_ranks) return fail(CL_ERR_INVALID, "cl_exchange_create: need 2 <= n_ranks <= 64 and 0 <= rank < n_ranks");
    if (env->wide) return fail(CL_ERR_UNSUPPORTED, "cl_exchange_create: building-tiled (wide) districts are not building-sharded across GPUs");
    if (env->d.central) return fail(CL_ERR_UNSUPPORTED, "cl_exchange_create: central-agent reward sums are not exchanged; use decentralised rewards");
    if (env->d.ch_n + env->d.wm_n > 0) return fail(CL_ERR_UNSUPPORTED, "cl_exchange_create: districts with electric vehicles / washing machines are not building-sharded");
    if (env->x_buf) return fail(CL_ERR_STATE, "cl_exchange_create: already created");
    if (env->kpi_unit) return fail(CL_ERR_UNSUPPORTED, "cl_exchange_create: online KPI accumulators need the whole district on one handle");
    // a block spins on its peers inside the step: every block of the launch must be resident (one wave), or blocks waiting for an SM
    // could be the ones a resident block waits for
    if (env->blocks > env->n_sm) return fail(CL_ERR_UNSUPPORTED, "cl_exchange_create: the launch must fit one wave (at most one block per SM); use fewer envs per GPU");
    const size_t bytes = exchange_bytes(env, n_ranks);
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->x_buf), bytes));
    CUDA_TRY(cudaMemset(env->x_buf, 0, bytes));              // epoch 0 never matches a step (epochs start at 1)
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->x_peers_dev), sizeof(uint2*) * n_ranks));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->x_err_dev), sizeof(int32_t)));
    CUDA_TRY(cudaMemset(env->x_err_dev, 0, sizeof(int32_t)));
    CUDA_TRY(cudaDeviceSynchronize());
    if (ipc_handle_out) {
        cudaIpcMemHandle_t h;
        CUDA_TRY(cudaIpcGetMemHandle(&h, env->x_buf));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::memcpy(ipc_handle_out, &h, sizeof(h));
    }
    if (buffer_out) *buffer_out = env->x_buf;
    env->d.x_rank = rank;
    env->d.x_n = -n_ranks;                                   // negative: created, not connected yet
    return CL_OK;
}

static int exchange_finish(cl_env* env, const std::vector<uint2*>& peers) {
    CUDA_TRY(cudaMemcpy(env->x_peers_dev, peers.data(), sizeof(uint2*) * peers.size(), cudaMemcpyHostToDevice));
    env->d.x_peers = env->x_peers_dev;
    env->d.x_err = env->x_err_dev;
    env->d.x_n = (int)peers.size();
    env->x_epoch = 0;
    return CL_OK;
}

extern "C" int cl_exchange_connect(cl_env* env, const void* ipc_handles) {
    if (!env || !ipc_handles) return fail(CL_ERR_INVALID, "cl_exchange_connect: null argument");
    if (env->d.x_n >= 0 || !env->x_buf) return fail(CL_ERR_STATE, "cl_exchange_connect: call cl_exchange_create first (once)");
    const int n = -env->d.x_n;
    std::vector<uint2*> peers((size_t)n, nullptr);
    for (int r = 0; r < n; ++r) {
        if (r == env->d.x_rank) { peers[(size_t)r] = env->x_buf; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(ipc_handles) + (size_t)r * sizeof(h), sizeof(h));
        void* q = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
        env->x_opened.push_back(q);
        peers[(size_t)r] = static_cast<uint2*>(q);
    }
    return exchange_finish(env, peers);
}

extern "C" int cl_exchange_connect_ptrs(cl_env* env, void* const* buffers, const int32_t* devices) {
    if (!env || !buffers) return fail(CL_ERR_INVALID, "cl_exchange_connect_ptrs: null argument");
    if (env->d.x_n >= 0 || !env->x_buf) return fail(CL_ERR_STATE, "cl_exchange_connect_ptrs: call cl_exchange_create first (once)");
    const int n = -env->d.x_n;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::vector<uint2*> peers((size_t)n, nullptr);
    for (int r = 0; r < n; ++r) {
        peers[(size_t)r] = static_cast<uint2*>(buffers[r]);
        if (r == env->d.x_rank) { if (buffers[r] != env->x_buf) return fail(CL_ERR_INVALID, "cl_exchange_connect_ptrs: buffers[rank] must be this handle's own buffer"); continue; }
        if (devices && devices[r] != dev) {
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, dev, devices[r]));
            if (!can) return fail(CL_ERR_UNSUPPORTED, "cl_exchange_connect_ptrs: no peer access between the devices");
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(CL_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    return exchange_finish(env, peers);
}

extern "C" int cl_exchange_status(cl_env* env, int32_t* timeouts, uint32_t* epoch) {
    if (!env) return fail(CL_ERR_INVALID, "cl_exchange_status: null env");
    if (env->d.x_n < 2) return fail(CL_ERR_STATE, "cl_exchange_status: no exchange connected");
    if (timeouts) { CUDA_TRY(cudaDeviceSynchronize()); CUDA_TRY(cudaMemcpy(timeouts, env->x_err_dev, sizeof(int32_t), cudaMemcpyDeviceToHost)); }
    if (epoch) *epoch = env->x_epoch;
    return CL_OK;
}

extern "C" int cl_obs_rows(cl_env* env, int32_t first_time_step, int32_t n_rows, float* rows, cl_stream stream) {
    if (!env || !rows) return fail(CL_ERR_INVALID, "cl_obs_rows: null argument");
    if (env->t < 0) return fail(CL_ERR_STATE, "cl_obs_rows: call cl_reset first");
    { const int rc = refresh_time(env, static_cast<cudaStream_t>(stream)); if (rc) return rc; }
    if (!env->d.stale || !env->d.uniform_start) return fail(CL_ERR_STATE, "cl_obs_rows: observation rows are only env-independent with stale_observations and one episode window for all envs");
    if (n_rows < 1 || first_time_step < 1 || first_time_step + n_rows > env->T) return fail(CL_ERR_INVALID, "cl_obs_rows: time steps must lie in [1, T - 1] (the observation at t = 0 differs per env)");
    const long total = (long)n_rows * env->d.L;
    obs_rows_kernel<<<(unsigned)std::min<long>((total + 255) / 256, 4096), 256, 0, static_cast<cudaStream_t>(stream)>>>(env->d, first_time_step, n_rows, rows);
    CUDA_TRY(cudaGetLastError());
    env->launches++;
    return CL_OK;
}

extern "C" int cl_ev_read(cl_env* env, float* soc_prev_dev, float* soc_dev, cl_stream stream) {
    if (!env) return fail(CL_ERR_INVALID, "cl_ev_read: null env");
    if (env->d.ev_n < 1) return fail(CL_ERR_STATE, "cl_ev_read: the district has no electric vehicles");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)env->d.E * env->d.ev_n;
    if (soc_prev_dev) CUDA_TRY(cudaMemcpyAsync(soc_prev_dev, env->d.ev_sf, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (soc_dev) CUDA_TRY(cudaMemcpyAsync(soc_dev, env->d.ev_sf + n, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return CL_OK;
}

extern "C" int cl_time_step(const cl_env* env, int32_t* t) {
    if (!env || !t) return fail(CL_ERR_INVALID, "cl_time_step: null argument");
    if (env->device_time && env->t >= 0) {
        // the counter may have been advanced by graph replays the host never saw
        int32_t v = 0;
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaMemcpy(&v, env->t_dev, sizeof(v), cudaMemcpyDeviceToHost));
        const_cast<cl_env*>(env)->t = v;
    }
    *t = env->t;
    return CL_OK;
}

extern "C" int cl_state_size(const cl_env* env, size_t* bytes) {
    if (!env || !bytes) return fail(CL_ERR_INVALID, "cl_state_size: null argument");
    *bytes = env->st_floats * sizeof(float) + env->dst_doubles * sizeof(double) + env->lst_floats * sizeof(float)
             + env->ev_sd_doubles * sizeof(double) + env->ev_sf_floats * sizeof(float) + env->ev_flag_bytes + env->wm_flag_bytes;
    return CL_OK;
}

extern "C" int cl_get_state(cl_env* env, void* dst_dev, cl_stream stream) {
    if (!env || !dst_dev) return fail(CL_ERR_INVALID, "cl_get_state: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* p = static_cast<char*>(dst_dev);
    if (env->dst_doubles) {   // doubles first keeps them 8-byte aligned inside the blob
        CUDA_TRY(cudaMemcpyAsync(p, env->d.dst, env->dst_doubles * sizeof(double), cudaMemcpyDeviceToDevice, st));
        p += env->dst_doubles * sizeof(double);
    }
    CUDA_TRY(cudaMemcpyAsync(p, env->d.st, env->st_floats * sizeof(float), cudaMemcpyDeviceToDevice, st));
    p += env->st_floats * sizeof(float);
    if (env->lst_floats) CUDA_TRY(cudaMemcpyAsync(p, env->d.lst, env->lst_floats * sizeof(float), cudaMemcpyDeviceToDevice, st));
    p += env->lst_floats * sizeof(float);
    if (env->ev_sf_floats) {      // vehicles / washing machines (the blob layout keeps 8-byte alignment: lst_floats and st_floats are even)
        CUDA_TRY(cudaMemcpyAsync(p, env->d.ev_sd, env->ev_sd_doubles * sizeof(double), cudaMemcpyDeviceToDevice, st)); p += env->ev_sd_doubles * sizeof(double);
        CUDA_TRY(cudaMemcpyAsync(p, env->d.ev_sf, env->ev_sf_floats * sizeof(float), cudaMemcpyDeviceToDevice, st)); p += env->ev_sf_floats * sizeof(float);
        CUDA_TRY(cudaMemcpyAsync(p, env->d.ev_flag, env->ev_flag_bytes, cudaMemcpyDeviceToDevice, st)); p += env->ev_flag_bytes;
        CUDA_TRY(cudaMemcpyAsync(p, env->d.wm_flag, env->wm_flag_bytes, cudaMemcpyDeviceToDevice, st));
    }
    return CL_OK;
}

extern "C" int cl_set_state(cl_env* env, const void* src_dev, int32_t time_step, cl_stream stream) {
    if (!env || !src_dev) return fail(CL_ERR_INVALID, "cl_set_state: null argument");
    if (env->T < 2 || time_step < 0 || time_step > env->T - 1) return fail(CL_ERR_STATE, "cl_set_state: time step outside the current episode (reset first)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const char* p = static_cast<const char*>(src_dev);
    if (env->dst_doubles) {
        CUDA_TRY(cudaMemcpyAsync(env->d.dst, p, env->dst_doubles * sizeof(double), cudaMemcpyDeviceToDevice, st));
        p += env->dst_doubles * sizeof(double);
    }
    CUDA_TRY(cudaMemcpyAsync(env->d.st, p, env->st_floats * sizeof(float), cudaMemcpyDeviceToDevice, st));
    p += env->st_floats * sizeof(float);
    if (env->lst_floats) CUDA_TRY(cudaMemcpyAsync(env->d.lst, p, env->lst_floats * sizeof(float), cudaMemcpyDeviceToDevice, st));
    p += env->lst_floats * sizeof(float);
    if (env->ev_sf_floats) {
        CUDA_TRY(cudaMemcpyAsync(env->d.ev_sd, p, env->ev_sd_doubles * sizeof(double), cudaMemcpyDeviceToDevice, st)); p += env->ev_sd_doubles * sizeof(double);
        CUDA_TRY(cudaMemcpyAsync(env->d.ev_sf, p, env->ev_sf_floats * sizeof(float), cudaMemcpyDeviceToDevice, st)); p += env->ev_sf_floats * sizeof(float);
        CUDA_TRY(cudaMemcpyAsync(env->d.ev_flag, p, env->ev_flag_bytes, cudaMemcpyDeviceToDevice, st)); p += env->ev_flag_bytes;
        CUDA_TRY(cudaMemcpyAsync(env->d.wm_flag, p, env->wm_flag_bytes, cudaMemcpyDeviceToDevice, st));
    }
    env->t = time_step;
    return publish_time(env, st);
}

extern "C" int cl_set_transforms(cl_env* env, const cl_obs_transform* obs_transform, const float* action_range, const float* action_low) {
    if (!env) return fail(CL_ERR_INVALID, "cl_set_transforms: null env");
    if ((action_range == nullptr) != (action_low == nullptr)) return fail(CL_ERR_INVALID, "cl_set_transforms: action_range and action_low go together");
    Dev& d = env->d;
    CUDA_TRY(cudaDeviceSynchronize());            // configuration call: no launch may still read the previous arrays
    d.obs_t = nullptr; d.act_range = nullptr; d.act_low = nullptr;
    if (obs_transform) {
        for (int k = 0; k < d.L; ++k) {
            const int fn = obs_transform[k].fn;
            if (fn != CL_OBS_FN_IDENTITY && fn != CL_OBS_FN_SIN && fn != CL_OBS_FN_COS) return fail(CL_ERR_INVALID, "cl_set_transforms: unknown observation function");
        }
        cl_obs_transform* p = nullptr;
        int rc = dev_copy(env, obs_transform, (size_t)d.L, &p);
        if (rc) return rc;
        d.obs_t = p;
    }
    if (action_range) {
        float *r = nullptr, *l = nullptr;
        int rc = dev_copy(env, action_range, (size_t)d.A, &r);
        if (!rc) rc = dev_copy(env, action_low, (size_t)d.A, &l);
        if (rc) return rc;
        d.act_range = r; d.act_low = l;
    }
    return build_obs_table(env);      // the table holds transformed values
}

extern "C" int cl_kpi_enable(cl_env* env, int32_t enable) {
    if (!env) return fail(CL_ERR_INVALID, "cl_kpi_enable: null env");
    if (enable && env->dynamics) return fail(CL_ERR_UNSUPPORTED, "cl_kpi_enable: the `_without_storage` baseline does not apply to LSTM-dynamics districts (use evaluate() on a recorded history)");
    if (enable && env->d.x_n != 0) return fail(CL_ERR_UNSUPPORTED, "cl_kpi_enable: a building-sharded handle sees only its own buildings' baseline");
    if (enable && (env->d.ch_n + env->d.wm_n) > 0) return fail(CL_ERR_UNSUPPORTED, "cl_kpi_enable: KPI accumulators do not cover districts with electric vehicles / washing machines");
    if (!enable) {
        CUDA_TRY(cudaDeviceSynchronize());
        if (env->kpi_unit) cudaFree(env->kpi_unit);
        if (env->kpi_env) cudaFree(env->kpi_env);
        env->kpi_unit = env->kpi_env = nullptr;
        env->kpi_fused = false;
        env->d.kpi_unit = env->d.kpi_env = nullptr; env->d.kpi_smem = 0;
        return CL_OK;
    }
    if (!env->kpi_unit) {
        const size_t nu = (size_t)env->d.U * CL_NKPI_UNIT, ne = (size_t)env->d.E * 2 * CL_NKPI_ENV;
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->kpi_unit), nu * sizeof(double)));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&env->kpi_env), ne * sizeof(double)));
        CUDA_TRY(cudaMemset(env->kpi_unit, 0, nu * sizeof(double)));
        CUDA_TRY(cudaMemset(env->kpi_env, 0, ne * sizeof(double)));
    }
    // fused into the step kernel when the district is a plain one (whole envs per block) and the accumulators fit its shared memory
    env->kpi_fused = false;
    env->d.kpi_smem = 0; env->d.kpi_unit = nullptr; env->d.kpi_env = nullptr;
    if (!env->wide && std::getenv("CL_B200_KPI_UNFUSED") == nullptr) {
        Dev t = env->d; t.kpi_smem = 1;
        const size_t smem = smem_bytes(t, env->threads + 32, true, env->precision == CL_PRECISION_FP64 ? 8 : 4);
        if (smem <= 200 * 1024) {
            ensure_smem_optin(smem);
            CUDA_TRY(cudaGetLastError());
            env->d.kpi_smem = 1; env->d.kpi_unit = env->kpi_unit; env->d.kpi_env = env->kpi_env;
            env->kpi_fused = true;
        }
    }
    return CL_OK;
}

extern "C" int cl_kpi_fused(const cl_env* env, int32_t* fused) {
    if (!env || !fused) return fail(CL_ERR_INVALID, "cl_kpi_fused: null argument");
    *fused = env->kpi_fused ? 1 : 0;
    return CL_OK;
}

extern "C" int cl_kpi_accumulate(cl_env* env, const float* trace, const float* district, cl_stream stream) {
    if (!env || !trace || !district) return fail(CL_ERR_INVALID, "cl_kpi_accumulate: null argument");
    if (!env->kpi_unit) return fail(CL_ERR_STATE, "cl_kpi_accumulate: call cl_kpi_enable first");
    if (env->kpi_fused) return CL_OK;                // already accumulated inside the step kernel
    { const int rc = refresh_time(env, static_cast<cudaStream_t>(stream)); if (rc) return rc; }
    if (env->t < 1) return fail(CL_ERR_STATE, "cl_kpi_accumulate: no step has been taken since cl_reset");
    int threads = ((std::min(env->d.B, 256) + 31) / 32) * 32;
    kpi_accumulate_kernel<<<env->d.E, threads, 0, static_cast<cudaStream_t>(stream)>>>(env->d, env->t - 1, trace, district, env->kpi_unit, env->kpi_env);
    CUDA_TRY(cudaGetLastError());
    env->launches++;
    return CL_OK;
}

extern "C" int cl_kpi_read(cl_env* env, double* unit_dev, double* env_dev, cl_stream stream) {
    if (!env) return fail(CL_ERR_INVALID, "cl_kpi_read: null env");
    if (!env->kpi_unit) return fail(CL_ERR_STATE, "cl_kpi_read: call cl_kpi_enable first");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (unit_dev) CUDA_TRY(cudaMemcpyAsync(unit_dev, env->kpi_unit, (size_t)env->d.U * CL_NKPI_UNIT * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (env_dev) CUDA_TRY(cudaMemcpyAsync(env_dev, env->kpi_env, (size_t)env->d.E * 2 * CL_NKPI_ENV * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return CL_OK;
}

// FP32 FMA throughput of the device (the roofline of the LSTM-dynamics path): 16 independent FFMA chains per thread
__global__ void fma_peak_kernel(float* __restrict__ out, int iters) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 1.0f + 1e-3f * (float)(threadIdx.x + j);
    const float b = 0.99999f, c = 1e-6f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], b, c);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) sum += a[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

extern "C" int cl_measure_fma_peak(double* tflops) {
    if (!tflops) return fail(CL_ERR_INVALID, "cl_measure_fma_peak: null argument");
    int dev = 0, n_sm = 148;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = n_sm * 8, threads = 256, iters = 2048;
    float* out = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&out), (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 16 * 8 * (double)iters * blocks * threads;
        if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    CUDA_TRY(cudaGetLastError());
    *tflops = best;
    return CL_OK;
}

// resident blocks per SM of the step kernel this handle launches (occupancy calculator with the launch's block size and dynamic
// shared memory): diagnostics for bench.py / profiles
extern "C" int cl_launch_occupancy(const cl_env* env, int32_t* blocks_per_sm, int32_t* smem_bytes_per_block) {
    if (!env) return fail(CL_ERR_INVALID, "cl_launch_occupancy: null env");
    const int nthreads = env->threads + 32;
    const bool f64 = env->precision == CL_PRECISION_FP64;
    size_t smem = smem_bytes(env->d, nthreads, false, f64 ? 8 : 4);
    if (env->blocks <= env->n_sm && smem < 116 * 1024) smem = 116 * 1024;     // launch_advance: a single-wave launch owns its SMs
    const void* fn = nullptr;
    if (env->wide) {
        fn = f64 ? (env->thermal ? (const void*)advance_kernel<double, true, false, 512, true> : (const void*)advance_kernel<double, false, false, 512, true>)
                 : (env->thermal ? (const void*)advance_kernel<float, true, false, 512, true> : (const void*)advance_kernel<float, false, false, 512, true>);
    } else if (nthreads <= (env->dynamics ? kDynMaxT : 512)) {
        fn = f64 ? (env->dynamics ? (const void*)advance_kernel<double, true, true, kDynMaxT> : (env->thermal ? (const void*)advance_kernel<double, true, false, 512> : (const void*)advance_kernel<double, false, false, 512>))
                 : (env->dynamics ? (const void*)advance_kernel<float, true, true, kDynMaxT> : (env->thermal ? (const void*)advance_kernel<float, true, false, 512> : (const void*)advance_kernel<float, false, false, 512>));
    } else {
        fn = f64 ? (env->dynamics ? (const void*)advance_kernel<double, true, true, 1024> : (env->thermal ? (const void*)advance_kernel<double, true, false, 1024> : (const void*)advance_kernel<double, false, false, 1024>))
                 : (env->dynamics ? (const void*)advance_kernel<float, true, true, 1024> : (env->thermal ? (const void*)advance_kernel<float, true, false, 1024> : (const void*)advance_kernel<float, false, false, 1024>));
    }
    int n = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, nthreads, smem));
    if (blocks_per_sm) *blocks_per_sm = n;
    if (smem_bytes_per_block) *smem_bytes_per_block = (int32_t)smem;
    return CL_OK;
}

extern "C" int cl_launch_geometry(const cl_env* env, int32_t* blocks, int32_t* threads, int32_t* tiles) {
    if (!env) return fail(CL_ERR_INVALID, "cl_launch_geometry: null env");
    if (blocks) *blocks = env->blocks;
    if (threads) *threads = env->threads + 32;
    if (tiles) *tiles = env->d.tiles;
    return CL_OK;
}

extern "C" int cl_launch_count(const cl_env* env, int64_t* n) {
    if (!env || !n) return fail(CL_ERR_INVALID, "cl_launch_count: null argument");
    *n = env->launches;
    return CL_OK;
}
