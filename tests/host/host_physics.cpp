// host_physics.cpp - TEST INFRASTRUCTURE ONLY.
// Compiles citylearn_b200/csrc/unit_physics.cuh (the exact per-unit code the CUDA kernels run) for the host so that
// `pytest -m "not gpu"` can check the physics against the oracle / golden traces without a GPU.  The product never
// loads this library; citylearn_b200 only ever calls libcitylearn_b200.so (CUDA) and fails loudly without it.
#include <cstring>
#include <vector>

#include "../../citylearn_b200/csrc/unit_physics.cuh"

using namespace cl;

template <typename R>
static void load_params_host(const double* P, const int32_t* ip, int B, int b, BuildingParams<R>& p) {
#define LD(k) ((R)(sizeof(R) == 4 ? (double)(float)P[(k) * B + b] : P[(k) * B + b]))
    p.bat_capacity = LD(CL_P_BAT_CAPACITY); p.bat_pnom = LD(CL_P_BAT_NOMINAL_POWER); p.bat_loss = LD(CL_P_BAT_LOSS);
    p.bat_clc = LD(CL_P_BAT_CLC); p.bat_dod = LD(CL_P_BAT_DOD);
    p.ratio = LD(CL_P_TIME_STEP_RATIO); p.hours = LD(CL_P_HOURS_PER_STEP);
    p.flags = ip[CL_IP_FLAGS * B + b]; p.pe_n = ip[CL_IP_PE_N * B + b]; p.cp_n = ip[CL_IP_CP_N * B + b];
    p.cd_pnom = LD(CL_P_CD_NOMINAL_POWER); p.cd_cop_num = LD(CL_P_CD_COP_NUM); p.cd_target = LD(CL_P_CD_TARGET);
    p.hd_pnom = LD(CL_P_HD_NOMINAL_POWER); p.hd_cop_num = LD(CL_P_HD_COP_NUM); p.hd_target = LD(CL_P_HD_TARGET); p.hd_eff = LD(CL_P_HD_EFFICIENCY);
    p.dd_pnom = LD(CL_P_DD_NOMINAL_POWER); p.dd_cop_num = LD(CL_P_DD_COP_NUM); p.dd_target = LD(CL_P_DD_TARGET); p.dd_eff = LD(CL_P_DD_EFFICIENCY);
    TankParams<R>* tanks[3] = {&p.cs, &p.hs, &p.ds};
    const int base[3] = {CL_P_CS_CAPACITY, CL_P_HS_CAPACITY, CL_P_DS_CAPACITY};
    const int fin[3] = {CL_F_CS_HAS_MAX_IN, CL_F_HS_HAS_MAX_IN, CL_F_DS_HAS_MAX_IN};
    const int fout[3] = {CL_F_CS_HAS_MAX_OUT, CL_F_HS_HAS_MAX_OUT, CL_F_DS_HAS_MAX_OUT};
    for (int i = 0; i < 3; ++i) {
        tanks[i]->capacity = LD(base[i]); tanks[i]->efficiency = LD(base[i] + 1); tanks[i]->loss = LD(base[i] + 2);
        tanks[i]->max_in = LD(base[i] + 4); tanks[i]->max_out = LD(base[i] + 5);
        tanks[i]->has_max_in = p.flags & fin[i]; tanks[i]->has_max_out = p.flags & fout[i];
    }
#undef LD
    derive_params(p);
}

// state: double [6][U] (soc_b, cap_deg, rte_b = sqrt(efficiency), soc_cs, soc_hs, soc_ds); dyn out: float [U][CL_NDYN]
template <typename R>
static void step_impl(int B, int E, int W, const double* P, const int32_t* ip, const float* table, const int32_t* start, int t,
                      const float* outage, int T, const float* actions, int A, double* state, float* dyn_out,
                      const unsigned char* control_cool) {
    const int U = B * E;
    std::vector<R> curves((size_t)4 * CL_MAX_CURVE * B);
    for (int k = 0; k < 4 * CL_MAX_CURVE; ++k)
        for (int b = 0; b < B; ++b) curves[(size_t)k * B + b] = (R)(sizeof(R) == 4 ? (double)(float)P[(CL_P_PE_X0 + k) * B + b] : P[(CL_P_PE_X0 + k) * B + b]);
    for (int e = 0; e < E; ++e) {
        const float* row = table + (size_t)(start[e] + t) * W;
        for (int b = 0; b < B; ++b) {
            const int u = e * B + b;
            BuildingParams<R> p;
            load_params_host<R>(P, ip, B, b, p);
            UnitInputs<R> in;
            auto col = [&](int k) { return row[ip[k * B + b]]; };
            in.nsl = (R)col(CL_IP_C_NSL);
            const R pv = (R)(sizeof(R) == 4 ? (double)(float)P[CL_P_PV_NOMINAL_POWER * B + b] : P[CL_P_PV_NOMINAL_POWER * B + b]);
            in.solar = -dvd(pv * (R)col(CL_IP_C_SOLAR), (R)1000);
            in.price = (R)col(CL_IP_C_PRICE); in.carbon = (R)col(CL_IP_C_CARBON);
            in.dhw_demand = (R)col(CL_IP_C_DHW_DEMAND); in.cooling_demand = (R)col(CL_IP_C_COOLING_DEMAND);
            in.heating_demand = (R)col(CL_IP_C_HEATING_DEMAND); in.t_out = (R)col(CL_IP_C_T_OUT);
            in.hvac_mode = (int32_t)col(CL_IP_C_HVAC_MODE);
            in.outage = outage && (p.flags & CL_F_SIMULATE_OUTAGE) && outage[b * T + t] > 0.f;
            const float* a = actions + (size_t)e * A;
            auto slot = [&](int k) { return ip[k * B + b]; };
            in.a_es = slot(CL_IP_A_ELECTRICAL_STORAGE) >= 0 ? (R)a[slot(CL_IP_A_ELECTRICAL_STORAGE)] : (R)0;
            in.a_cooling_device = slot(CL_IP_A_COOLING_DEVICE) >= 0 ? (R)a[slot(CL_IP_A_COOLING_DEVICE)] : (R)NAN;
            in.a_heating_device = slot(CL_IP_A_HEATING_DEVICE) >= 0 ? (R)a[slot(CL_IP_A_HEATING_DEVICE)] : (R)NAN;
            in.a_cs = slot(CL_IP_A_COOLING_STORAGE) >= 0 ? (R)a[slot(CL_IP_A_COOLING_STORAGE)] : (R)0;
            in.a_hs = slot(CL_IP_A_HEATING_STORAGE) >= 0 ? (R)a[slot(CL_IP_A_HEATING_STORAGE)] : (R)0;
            in.a_ds = slot(CL_IP_A_DHW_STORAGE) >= 0 ? (R)a[slot(CL_IP_A_DHW_STORAGE)] : (R)0;
            in.control_cooling_demand = control_cool && control_cool[u];
            in.control_heating_demand = false;
            UnitState<R> s;
            s.soc_b = (R)state[0 * U + u]; s.cap_deg = (R)state[1 * U + u]; s.rte_b = (R)state[2 * U + u];
            s.soc_cs = (R)state[3 * U + u]; s.soc_hs = (R)state[4 * U + u]; s.soc_ds = (R)state[5 * U + u];
            UnitResult<R> o;
            unit_step<R, true>(p, StridedCurves<R, R>{curves.data() + b, B}, t, in, s, o);
            state[0 * U + u] = (double)s.soc_b; state[1 * U + u] = (double)s.cap_deg; state[2 * U + u] = (double)s.rte_b;
            state[3 * U + u] = (double)s.soc_cs; state[4 * U + u] = (double)s.soc_hs; state[5 * U + u] = (double)s.soc_ds;
            float* d = dyn_out + (size_t)u * CL_NDYN;
            d[CL_DYN_ELECTRICAL_STORAGE_SOC] = (float)s.soc_b; d[CL_DYN_COOLING_STORAGE_SOC] = (float)s.soc_cs;
            d[CL_DYN_HEATING_STORAGE_SOC] = (float)s.soc_hs; d[CL_DYN_DHW_STORAGE_SOC] = (float)s.soc_ds;
            d[CL_DYN_NET_ELECTRICITY_CONSUMPTION] = (float)o.net;
            d[CL_DYN_COOLING_DEMAND] = (float)(o.e_from_cool + fabs(rmin(o.eb_cs, (R)0)));
            d[CL_DYN_HEATING_DEMAND] = (float)(o.e_from_heat + fabs(rmin(o.eb_hs, (R)0)));
            d[CL_DYN_DHW_DEMAND] = (float)(o.e_from_dhw + fabs(rmin(o.eb_ds, (R)0)));
            d[CL_DYN_COOLING_ELECTRICITY_CONSUMPTION] = (float)(o.ec_cool * p.ratio);
            d[CL_DYN_HEATING_ELECTRICITY_CONSUMPTION] = (float)(o.ec_heat * p.ratio);
            d[CL_DYN_DHW_ELECTRICITY_CONSUMPTION] = (float)(o.ec_dhw * p.ratio);
            d[CL_DYN_ELECTRICAL_STORAGE_ELECTRICITY_CONSUMPTION] = (float)(o.ec_bat * p.ratio);
            d[CL_DYN_NON_SHIFTABLE_LOAD_ELECTRICITY_CONSUMPTION] = (float)(o.ec_nsl * p.ratio);
            d[CL_DYN_ELECTRICAL_STORAGE_ENERGY_BALANCE] = (float)o.eb_bat; d[CL_DYN_COOLING_STORAGE_ENERGY_BALANCE] = (float)o.eb_cs;
            d[CL_DYN_HEATING_STORAGE_ENERGY_BALANCE] = (float)o.eb_hs; d[CL_DYN_DHW_STORAGE_ENERGY_BALANCE] = (float)o.eb_ds;
            d[CL_DYN_NET_ELECTRICITY_CONSUMPTION_COST] = (float)o.cost; d[CL_DYN_NET_ELECTRICITY_CONSUMPTION_EMISSION] = (float)o.emission;
            d[CL_DYN_ELECTRICAL_STORAGE_DEGRADED_CAPACITY] = (float)s.cap_deg;
            d[CL_DYN_ENERGY_TO_NON_SHIFTABLE_LOAD] = (float)o.e_to_nsl; d[CL_DYN_COOLING_DEMAND_SERIES] = (float)o.cool_dem;
            d[CL_DYN_HEATING_DEMAND_SERIES] = (float)o.heat_dem;
        }
    }
}

extern "C" void host_step(int precision, int B, int E, int W, const double* P, const int32_t* ip, const float* table, const int32_t* start,
                          int t, const float* outage, int T, const float* actions, int A, double* state, float* dyn_out,
                          const unsigned char* control_cool) {
    if (precision == CL_PRECISION_FP64) step_impl<double>(B, E, W, P, ip, table, start, t, outage, T, actions, A, state, dyn_out, control_cool);
    else step_impl<float>(B, E, W, P, ip, table, start, t, outage, T, actions, A, state, dyn_out, control_cool);
}

// segment index chosen by the device-side uniform-grid search (SmemCurves with its index) and by the reference-order scan
// (StridedCurves) for the points x[0..nx): returns -1 when the curve has no index (two points in one cell), else the mismatches.
template <typename R>
static int curve_check_impl(const double* xs, const double* ys, int n, const double* x, int nx, int* first_bad) {
    R tab[kCurveTab];
    for (int k = 0; k < CL_MAX_CURVE; ++k) {
        tab[k] = k < n ? (R)xs[k] : Num<R>::inf(); tab[CL_MAX_CURVE + k] = k < n ? (R)ys[k] : (R)0;
        tab[2 * CL_MAX_CURVE + k] = tab[k]; tab[3 * CL_MAX_CURVE + k] = tab[CL_MAX_CURVE + k];
        tab[4 * CL_MAX_CURVE + k] = (R)0; tab[5 * CL_MAX_CURVE + k] = (R)0;
    }
    uint8_t lut[2 * kCurveLutStride];
    if (!build_curve_lut(xs, 1, n, sizeof(R) == 4, lut)) return -1;
    std::memcpy(lut + kCurveLutStride, lut, kCurveLutStride);
    R strided[4 * CL_MAX_CURVE];
    for (int k = 0; k < CL_MAX_CURVE; ++k) { strided[k] = (R)xs[k < n ? k : n - 1]; strided[CL_MAX_CURVE + k] = (R)ys[k < n ? k : n - 1]; }
    const SmemCurves<R> a{tab, n, lut};
    const SmemCurves<R> loop{tab, n, nullptr};
    const StridedCurves<R, R> b{strided, 1};
    int bad = 0;
    for (int i = 0; i < nx; ++i) {
        const CurveSegment<R> ga = a.segment(CL_CURVE_PE, n, (R)x[i]), gl = loop.segment(CL_CURVE_PE, n, (R)x[i]), gb = b.segment(CL_CURVE_PE, n, (R)x[i]);
        if (!(ga.x0 == gb.x0 && ga.y0 == gb.y0 && ga.dy == gb.dy && gl.x0 == gb.x0 && gl.y0 == gb.y0)) { if (!bad) *first_bad = i; ++bad; }
    }
    return bad;
}
extern "C" int host_curve_check(int precision, const double* xs, const double* ys, int n, const double* x, int nx, int* first_bad) {
    return precision == CL_PRECISION_FP64 ? curve_check_impl<double>(xs, ys, n, x, nx, first_bad) : curve_check_impl<float>(xs, ys, n, x, nx, first_bad);
}

// One charger update with the device's `charger_step` (+ the vehicle battery's `battery_charge`): ev = the vehicle's parameter row
// [CL_NPARAM], chp = the charger's [CL_NCHP]; state = {soc (soc[t-1] entry), degraded capacity, sqrt(efficiency of the last charge)}
// in / out; out = {electricity consumption, commanded kWh}.  Returns 1 when the battery was charged / discharged.
template <typename R>
static int charger_impl(const double* chp, double action, int connected, const double* ev, int pe_n, int cp_n, int first_charge, double* state, double* out) {
    ChargerParams<R> q;
    q.max_c = (R)chp[CL_CH_MAX_C]; q.min_c = (R)chp[CL_CH_MIN_C]; q.max_d = (R)chp[CL_CH_MAX_D]; q.min_d = (R)chp[CL_CH_MIN_D]; q.eff = (R)chp[CL_CH_EFF];
    q.c_n = (int)chp[CL_CH_C_N]; q.d_n = (int)chp[CL_CH_D_N]; q.curves = chp + CL_CH_C_X0;
    BuildingParams<R> p;
    p.bat_capacity = (R)ev[CL_P_BAT_CAPACITY]; p.bat_pnom = (R)ev[CL_P_BAT_NOMINAL_POWER]; p.bat_loss = (R)ev[CL_P_BAT_LOSS];
    p.bat_clc = (R)ev[CL_P_BAT_CLC]; p.bat_dod = (R)ev[CL_P_BAT_DOD]; p.ratio = (R)ev[CL_P_TIME_STEP_RATIO]; p.hours = (R)ev[CL_P_HOURS_PER_STEP];
    p.flags = 0; p.pe_n = pe_n; p.cp_n = cp_n;
    derive_params(p);
    R curves[4 * CL_MAX_CURVE];
    for (int k = 0; k < 4 * CL_MAX_CURVE; ++k) curves[k] = (R)ev[CL_P_PE_X0 + k];
    UnitState<R> s;
    s.soc_b = (R)state[0]; s.cap_deg = (R)state[1]; s.rte_b = (R)state[2]; s.soc_cs = s.soc_hs = s.soc_ds = (R)0;
    float kwh = 0.f; bool charged = false;
    const float ec = charger_step<R>(q, action, connected != 0, p, StridedCurves<R, R>{curves, 1}, first_charge != 0, s, (double)p.hours, kwh, charged);
    state[0] = (double)s.soc_b; state[1] = (double)s.cap_deg; state[2] = (double)s.rte_b;
    out[0] = (double)ec; out[1] = (double)kwh;
    return charged ? 1 : 0;
}
extern "C" int host_charger_step(int precision, const double* chp, double action, int connected, const double* ev, int pe_n, int cp_n, int first_charge,
                                 double* state, double* out) {
    return precision == CL_PRECISION_FP64 ? charger_impl<double>(chp, action, connected, ev, pe_n, cp_n, first_charge, state, out)
                                          : charger_impl<float>(chp, action, connected, ev, pe_n, cp_n, first_charge, state, out);
}
