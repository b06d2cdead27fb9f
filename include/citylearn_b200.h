/*
 * citylearn_b200.h - C ABI of the B200-native CityLearn step path.
 *
 * The reference (intelligent-environments-lab/CityLearn v2.4.2) is pure Python and has no FFI; the entry
 * points below are what a binding for its hot path would call.  Each one names the reference
 * interface it replaces (file:line relative to the reference tree):
 *
 *   cl_create   <- CityLearnEnv.__init__ / _load ............ citylearn/citylearn.py:133-271, 1973-2170
 *                  (the schema is parsed on the host by citylearn_b200/schema.py; the descriptor carries
 *                   the flattened result: parameters, the time-series table, the observation layout)
 *   cl_reset    <- CityLearnEnv.reset ...................... citylearn/citylearn.py:1829-1886
 *                  Building.reset .......................... citylearn/building.py:2526-2564
 *   cl_step     <- CityLearnEnv.step ....................... citylearn/citylearn.py:978-1056
 *                  Building.apply_actions .................. citylearn/building.py:1500-1634
 *                  Building.update_variables ............... citylearn/building.py:2615-2703
 *                  CityLearnEnv.update_variables ........... citylearn/citylearn.py:1888-1918
 *                  RewardFunction.calculate (built-ins) .... citylearn/reward_function.py:65-386
 *                  CityLearnEnv.observations ............... citylearn/citylearn.py:451-485
 *   cl_rollout  <- the caller's `while not env.terminated: env.step(a)` loop
 *                  (citylearn/agents/base.py:155-176) for open-loop action blocks
 *   cl_get_state / cl_set_state <- (no reference equivalent; checkpoint / resume of the mutable state)
 *   cl_set_outage <- Building.reset_power_outage_signal ..... citylearn/building.py:2566-2594
 *
 * Conventions: every pointer marked "dev" is device memory owned by the caller (e.g. torch tensors'
 * data_ptr()); "host" pointers are only read during the call.  All entry points enqueue work on the
 * given CUDA stream and return without synchronising.  Return value: CL_OK or an error code; the
 * message of the last error of the calling thread is available from cl_last_error().
 * One cl_env must not be used from two streams at once; distinct handles are independent.
 */
#ifndef CITYLEARN_B200_H
#define CITYLEARN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CL_ABI_VERSION 2

typedef enum cl_status {
    CL_OK = 0,
    CL_ERR_INVALID = 1,      /* bad argument / descriptor */
    CL_ERR_CUDA = 2,         /* CUDA runtime error (message has the cudaError string) */
    CL_ERR_UNSUPPORTED = 3,  /* feature outside the accelerated path */
    CL_ERR_STATE = 4         /* call sequence error (e.g. step before reset, step after the episode end) */
} cl_status;

/* ---- per-building float parameters: column k of params[CL_NPARAM][B] (host, double) ------------- */
enum cl_building_param {
    CL_P_BAT_CAPACITY = 0, CL_P_BAT_NOMINAL_POWER, CL_P_BAT_EFFICIENCY0, CL_P_BAT_LOSS, CL_P_BAT_CLC, CL_P_BAT_DOD,
    CL_P_BAT_INITIAL_SOC,
    CL_P_CS_CAPACITY, CL_P_CS_EFFICIENCY, CL_P_CS_LOSS, CL_P_CS_INITIAL_SOC, CL_P_CS_MAX_IN, CL_P_CS_MAX_OUT,
    CL_P_HS_CAPACITY, CL_P_HS_EFFICIENCY, CL_P_HS_LOSS, CL_P_HS_INITIAL_SOC, CL_P_HS_MAX_IN, CL_P_HS_MAX_OUT,
    CL_P_DS_CAPACITY, CL_P_DS_EFFICIENCY, CL_P_DS_LOSS, CL_P_DS_INITIAL_SOC, CL_P_DS_MAX_IN, CL_P_DS_MAX_OUT,
    CL_P_CD_NOMINAL_POWER, CL_P_CD_COP_NUM, CL_P_CD_TARGET,
    CL_P_HD_NOMINAL_POWER, CL_P_HD_COP_NUM, CL_P_HD_TARGET, CL_P_HD_EFFICIENCY,
    CL_P_DD_NOMINAL_POWER, CL_P_DD_COP_NUM, CL_P_DD_TARGET, CL_P_DD_EFFICIENCY,
    CL_P_TIME_STEP_RATIO, CL_P_HOURS_PER_STEP,
    CL_P_PE_X0, CL_P_PE_X1, CL_P_PE_X2, CL_P_PE_X3, CL_P_PE_X4, CL_P_PE_X5, CL_P_PE_X6, CL_P_PE_X7,
    CL_P_PE_Y0, CL_P_PE_Y1, CL_P_PE_Y2, CL_P_PE_Y3, CL_P_PE_Y4, CL_P_PE_Y5, CL_P_PE_Y6, CL_P_PE_Y7,
    CL_P_CP_X0, CL_P_CP_X1, CL_P_CP_X2, CL_P_CP_X3, CL_P_CP_X4, CL_P_CP_X5, CL_P_CP_X6, CL_P_CP_X7,
    CL_P_CP_Y0, CL_P_CP_Y1, CL_P_CP_Y2, CL_P_CP_Y3, CL_P_CP_Y4, CL_P_CP_Y5, CL_P_CP_Y6, CL_P_CP_Y7,
    CL_P_DYN_TIN_MIN, CL_P_DYN_TIN_MAX, CL_P_DYN_CDEM_MIN, CL_P_DYN_CDEM_MAX,
    CL_P_PV_NOMINAL_POWER,   /* C_SOLAR is the raw inverter series [W/kW]; solar_generation = -(P * s / 1000), building.py:2554 */
    CL_NPARAM
};
#define CL_MAX_CURVE 8

/* ---- per-building int parameters: column k of iparams[CL_NIPARAM][B] (host, int32) -------------- */
enum cl_building_iparam {
    CL_IP_FLAGS = 0, CL_IP_PE_N, CL_IP_CP_N,
    CL_IP_A_COOLING_DEVICE, CL_IP_A_HEATING_DEVICE, CL_IP_A_COOLING_OR_HEATING_DEVICE,
    CL_IP_A_COOLING_STORAGE, CL_IP_A_HEATING_STORAGE, CL_IP_A_DHW_STORAGE, CL_IP_A_ELECTRICAL_STORAGE,
    CL_IP_C_NSL, CL_IP_C_DHW_DEMAND, CL_IP_C_COOLING_DEMAND, CL_IP_C_HEATING_DEMAND, CL_IP_C_SOLAR, CL_IP_C_T_OUT,
    CL_IP_C_PRICE, CL_IP_C_CARBON, CL_IP_C_HVAC_MODE, CL_IP_C_T_IN, CL_IP_C_COOL_SP, CL_IP_C_HEAT_SP,
    CL_IP_C_COMFORT_BAND, CL_IP_C_OCCUPANT,
    CL_IP_DYN_C_INPUTS, CL_IP_DYN_N_INPUTS, CL_IP_DYN_SLOT_TIN, CL_IP_DYN_SLOT_CDEM, CL_IP_DYN_W_OFFSET,
    CL_IP_DYN_LOOKBACK, CL_IP_DYN_HIDDEN,
    CL_NIPARAM
};

/* CL_IP_FLAGS bits */
#define CL_F_HEATING_IS_HEAT_PUMP (1 << 0)
#define CL_F_DHW_IS_HEAT_PUMP     (1 << 1)
#define CL_F_SIMULATE_OUTAGE      (1 << 2)
#define CL_F_DYNAMICS             (1 << 3)
#define CL_F_HAS_THERMAL          (1 << 4)
#define CL_F_CS_HAS_MAX_IN        (1 << 5)
#define CL_F_CS_HAS_MAX_OUT       (1 << 6)
#define CL_F_HS_HAS_MAX_IN        (1 << 7)
#define CL_F_HS_HAS_MAX_OUT       (1 << 8)
#define CL_F_DS_HAS_MAX_IN        (1 << 9)
#define CL_F_DS_HAS_MAX_OUT       (1 << 10)
#define CL_F_CS_CAPACITY_F32      (1 << 11) /* autosized tank: np.float32 capacity -> float32 `action * capacity` (energy_model.py:770-795) */
#define CL_F_HS_CAPACITY_F32      (1 << 12)
#define CL_F_CD_NOMINAL_F32       (1 << 13) /* autosized heat pump: np.float32 nominal power -> float32 `action * nominal_power` (building.py:3110,3146) */
#define CL_F_HD_NOMINAL_F32       (1 << 14)

/* ---- per-unit dynamic values at time step t (trace output, DYN observation slots) --------------- */
enum cl_dyn {
    CL_DYN_ELECTRICAL_STORAGE_SOC = 0, CL_DYN_COOLING_STORAGE_SOC, CL_DYN_HEATING_STORAGE_SOC, CL_DYN_DHW_STORAGE_SOC,
    CL_DYN_NET_ELECTRICITY_CONSUMPTION, CL_DYN_COOLING_DEMAND, CL_DYN_HEATING_DEMAND, CL_DYN_DHW_DEMAND,
    CL_DYN_COOLING_ELECTRICITY_CONSUMPTION, CL_DYN_HEATING_ELECTRICITY_CONSUMPTION, CL_DYN_DHW_ELECTRICITY_CONSUMPTION,
    CL_DYN_COOLING_STORAGE_ELECTRICITY_CONSUMPTION, CL_DYN_HEATING_STORAGE_ELECTRICITY_CONSUMPTION,
    CL_DYN_DHW_STORAGE_ELECTRICITY_CONSUMPTION, CL_DYN_ELECTRICAL_STORAGE_ELECTRICITY_CONSUMPTION,
    CL_DYN_INDOOR_DRY_BULB_TEMPERATURE, CL_DYN_NON_SHIFTABLE_LOAD_ELECTRICITY_CONSUMPTION,
    CL_DYN_ELECTRICAL_STORAGE_ENERGY_BALANCE, CL_DYN_COOLING_STORAGE_ENERGY_BALANCE, CL_DYN_HEATING_STORAGE_ENERGY_BALANCE,
    CL_DYN_DHW_STORAGE_ENERGY_BALANCE, CL_DYN_NET_ELECTRICITY_CONSUMPTION_COST, CL_DYN_NET_ELECTRICITY_CONSUMPTION_EMISSION,
    CL_DYN_ELECTRICAL_STORAGE_DEGRADED_CAPACITY,
    /* series needed by the KPI table (CityLearnEnv.evaluate, citylearn.py:1136-1323) */
    CL_DYN_ENERGY_TO_NON_SHIFTABLE_LOAD, CL_DYN_COOLING_DEMAND_SERIES, CL_DYN_HEATING_DEMAND_SERIES,
    CL_NDYN
};

/* ---- observation descriptor: obs_desc[L][4] = (kind, a, b, building) ----------------------------- */
enum cl_obs_kind {
    CL_OBS_TS = 0,          /* a = table column: value of the series at the observed time step            */
    CL_OBS_DYN = 1,         /* a = cl_dyn slot of `building`; zero after a step in reference-parity mode  */
    CL_OBS_OUTAGE = 2,      /* power-outage signal of `building` at the observed time step                */
    CL_OBS_TS_MINUS_TS = 3, /* reserved                                                                    */
    CL_OBS_STATE = 4        /* a = slot of the env's charging-constraint state (cl_ev_desc.cc_*): k * CL_CC_SLOTS + (0: building headroom kW;
                               1..CL_MAX_PHASES: phase headroom kW; CL_CC_SLOTS - 1: violation kWh) of the last applied actions */
};

/* ---- built-in reward functions (citylearn/reward_function.py) ------------------------------------ */
enum cl_reward_id {
    CL_REWARD_DEFAULT = 0,            /* RewardFunction :65-88   p[0] = exponent                          */
    CL_REWARD_MARL = 1,               /* MARL :132-143                                                     */
    CL_REWARD_INDEPENDENT_SAC = 2,    /* IndependentSACReward :159-168                                     */
    CL_REWARD_SOLAR_PENALTY = 3,      /* SolarPenaltyReward :189-214                                       */
    CL_REWARD_COMFORT = 4,            /* ComfortReward :269-334  p[0]=band (NaN: series) p[1]=lower p[2]=higher exponent */
    CL_REWARD_SOLAR_PENALTY_AND_COMFORT = 5, /* :381-386         p[3], p[4] = coefficients                 */
    CL_REWARD_ELECTRIC_VEHICLES = 6,  /* Electric_Vehicles_Reward_Function :389-523 (default weights; needs cl_ev_desc) p[0] = charging-constraint penalty coefficient */
    CL_REWARD_NONE = -1               /* rewards are computed by the caller from the trace (custom RewardFunction) */
};

enum cl_precision {
    CL_PRECISION_FP32 = 0,  /* float arithmetic (north-star contract: <= 1e-5 scaled-relative of the reference)      */
    CL_PRECISION_FP64 = 1   /* the reference's own float64-intermediate / float32-storage flow (bit-exact physics)    */
};

/*
 * Electric vehicles, chargers and washing machines (citylearn/electric_vehicle_charger.py, electric_vehicle.py,
 * energy_model.py:1244-1398, citylearn.py:1325-1475).  Which vehicle is plugged in where, arrival SOCs and the SOC drift of away
 * vehicles do not depend on the actions: the host compiles them into table columns (citylearn_b200/ev.py) and the device keeps, per
 * (vehicle, env), the soc[t-1] / soc[t] entries, the degraded capacity, the last round-trip efficiency and a "has charged" flag.
 * All arrays are HOST arrays copied by cl_create.  Districts with vehicles use whole-env blocks (no building tiles), one episode
 * window for all envs and reference-parity (stale) observations.
 */
enum cl_charger_param {
    CL_CH_MAX_C = 0, CL_CH_MIN_C, CL_CH_MAX_D, CL_CH_MIN_D, CL_CH_EFF, CL_CH_C_N, CL_CH_D_N,
    CL_CH_C_X0, CL_CH_C_Y0 = CL_CH_C_X0 + 8, CL_CH_D_X0 = CL_CH_C_Y0 + 8, CL_CH_D_Y0 = CL_CH_D_X0 + 8, CL_NCHP = CL_CH_D_Y0 + 8
};
#define CL_MAX_CHARGERS_PER_BUILDING 4
#define CL_MAX_PHASES 4
#define CL_CC_SLOTS (2 + CL_MAX_PHASES)
typedef struct cl_ev_desc {
    int32_t n_ev, n_chargers, n_machines;
    const double* ev_params;      /* [n_ev][CL_NPARAM]: the CL_P_BAT_* / curve / time-step entries of every vehicle battery      */
    const int32_t* ev_iparams;    /* [n_ev][2]: points of the power-efficiency and capacity-power curves                          */
    const int32_t* ev_cols;       /* [n_ev][4] table columns: association SOC, pre-connection SOC, (unused), SOC at an episode start; NaN = none */
    const double* ev_drift;       /* [n_rows][n_ev] factor on soc[t-1] for a vehicle that is away on that row, NaN = none (float64) */
    const int32_t* ch_building;   /* [n_chargers] ascending                                                                        */
    const int32_t* ch_action;     /* [n_chargers] slot in the district action vector or -1                                         */
    const int32_t* ch_cols;       /* [n_chargers][4] table columns: connected (0/1), vehicle index, required SOC, hours to departure */
    const double* ch_params;      /* [n_chargers][CL_NCHP]                                                                         */
    const int32_t* wm_building;   /* [n_machines] ascending                                                                        */
    const int32_t* wm_action;     /* [n_machines]                                                                                  */
    const int32_t* wm_cols;       /* [n_machines][4] table columns: window start, window end, load (sum of the profile), profile length.
                                     The load column is followed by max(profile length) - 1 columns holding the sum of the first 1, 2, ...
                                     entries: a cycle whose profile would run past the episode end adds only those (energy_model.py:1325-1327) */
    /* charging constraints (citylearn/building.py:764-989): caps on the sum of a building's positive charger requests and on the
       chargers of each phase; the step scales the charger actions down, keeps headroom / violation as per-env state (CL_OBS_STATE,
       not touched by cl_reset - the reference does not reset it either) and Electric_Vehicles_Reward_Function subtracts
       violation x reward_params[0].  May be 0 / NULL. */
    int32_t n_constrained;
    const int32_t* cc_building;   /* [n_constrained] ascending                                                                      */
    const double* cc_limits;      /* [n_constrained][1 + CL_MAX_PHASES] kW: building cap, phase caps; NaN = none                    */
    const int32_t* cc_members;    /* [n_constrained][CL_MAX_PHASES][CL_MAX_CHARGERS_PER_BUILDING] index of a phase's chargers within
                                     the building's chargers, in the phase's list order, -1 = end                                   */
    const int32_t* cc_flags;      /* [n_constrained] bit 0: the violation is exposed (observation, reward penalty)                  */
} cl_ev_desc;

typedef struct cl_district_desc {
    int32_t abi_version;         /* CL_ABI_VERSION */
    int32_t n_buildings;         /* B */
    int32_t n_envs;              /* E parallel environments held by this handle (this GPU's shard) */
    int32_t n_rows;              /* rows of the time-series table (dataset length) */
    int32_t n_cols;              /* W columns per row */
    int32_t action_dim;          /* sum of active actions over buildings */
    int32_t obs_dim;             /* L values per env observation row */
    int32_t central_agent;       /* 1: reward is the district sum [E,1]; 0: [E,B] */
    int32_t reward_id;           /* cl_reward_id */
    int32_t precision;           /* cl_precision */
    int32_t stale_observations;  /* 1: reference parity (DYN observations read 0 after a step) */
    int32_t lstm_weight_count;   /* floats in lstm_weights (0: no dynamics) */
    double reward_params[8];
    const float* table;          /* host [n_rows][n_cols] */
    const double* params;        /* host [CL_NPARAM][B] */
    const int32_t* iparams;      /* host [CL_NIPARAM][B] */
    const int32_t* obs_desc;     /* host [L][4] */
    const float* lstm_weights;   /* host, per-building blocks at iparams[CL_IP_DYN_W_OFFSET] */
    const cl_ev_desc* ev;        /* electric vehicles / chargers / washing machines, or NULL */
} cl_district_desc;

typedef struct cl_env cl_env;    /* opaque */
typedef void* cl_stream;         /* cudaStream_t */

/* Build the device-side district: copies the table, parameters and layouts to the current CUDA device. */
int cl_create(const cl_district_desc* desc, cl_env** out);
int cl_destroy(cl_env* env);

/* Power-outage signals of the coming episode: host [B][episode_time_steps] float 0/1 (NULL: no outages). */
int cl_set_outage(cl_env* env, const float* signals, int32_t episode_time_steps, cl_stream stream);

/*
 * Start an episode.  episode_start: dev [E] int32 table row of time step 0 of every env, or NULL with
 * `uniform_start` used for all envs.  episode_time_steps: T (the episode has T-1 steps, citylearn.py:372-376).
 * obs: dev [E][L] observation at t = 0 (may be NULL).
 */
int cl_reset(cl_env* env, const int32_t* episode_start, int32_t uniform_start, int32_t episode_time_steps,
             float* obs, cl_stream stream);

/*
 * Advance every env by one time step.
 *   actions  dev [E][action_dim]      (district action vector: buildings in order, active actions in schema order)
 *   obs      dev [E][L]               observation at t+1                        (NULL: not materialised)
 *   reward   dev [E][B] or [E][1]     reward of step t                          (NULL allowed)
 *   district dev [E][3]               sum over buildings of net, cost, emission (NULL allowed)
 *   trace    dev [E][B][CL_NDYN]      per-unit values at t                      (NULL allowed)
 */
int cl_step(cl_env* env, const float* actions, float* obs, float* reward, float* district, float* trace,
            cl_stream stream);

/*
 * K consecutive steps in ONE launch with pre-resident actions (open-loop block / on-device policy output).
 *   actions dev [K][E][action_dim]; obs dev [K][E][L]; reward dev [K][E][R]; district dev [K][E][3] (NULL allowed each
 *   except actions).  Equivalent to K calls of cl_step.
 */
int cl_rollout(cl_env* env, int32_t n_steps, const float* actions, float* obs, float* reward, float* district,
               cl_stream stream);

/*
 * Device-resident time step: closed loops without the host (policy -> step -> policy ..., citylearn/agents/base.py:155-176)
 * captured ONCE as a CUDA graph and replayed.  cl_step / cl_rollout take the time step from the host handle, so a captured launch
 * would repeat the same step; cl_advance_device instead reads it from a counter on the device and advances the counter in-stream
 * (a second, one-thread kernel), so the captured sequence is replayable.  A launch that would run past the end of the episode is a
 * no-op (the counter stays at T - 1).
 *   cl_device_time_enable  once per handle, OUTSIDE any capture: publishes the host's time step to the counter; from then on the
 *                          counter is authoritative - the host entry points read it back (one stream synchronisation) when they
 *                          need the time step, and cl_reset / cl_set_state / cl_step / cl_rollout keep it up to date
 *   cl_advance_device      n_steps like cl_rollout (n_steps = 1: cl_step); capturable: no allocation, no synchronisation
 */
int cl_device_time_enable(cl_env* env, cl_stream stream);
int cl_advance_device(cl_env* env, int32_t n_steps, const float* actions, float* obs, float* reward, float* district, cl_stream stream);

/*
 * Building-sharded districts (SURVEY.md §8e "district all-reduce variant"; citylearn/citylearn.py:1908-1918 sums net / cost /
 * emission over ALL buildings, citylearn/reward_function.py:132-143 reads that sum).  Rank r of n holds a handle over ITS buildings of
 * every env (same n_envs, same episode windows, decentralised rewards); inside cl_step / cl_rollout the per-env district sums are
 * completed across the ranks by an all-gather of the partial sums through peer memory over NVLink: each (quantity, env) partial is
 * pushed as one 8-byte {value, epoch} store into every rank's slot array and summed in rank order, so `district` and the district
 * term of MARL are the same on every rank (float32, associated per rank: within 1e-5 of the single-GPU sums).
 * Every rank must issue the same sequence of cl_step / cl_rollout calls (same n_steps, `district` NULL or not on all ranks).
 *   cl_exchange_create        allocates this rank's slot array; ipc_handle_out (64 bytes, may be NULL) is its cudaIpcMemHandle for
 *                             ranks in OTHER processes, buffer_out (may be NULL) its device pointer for ranks in THIS process
 *   cl_exchange_connect       multi-process: ipc_handles[n_ranks][64], gathered from all ranks (entry `rank` is ignored)
 *   cl_exchange_connect_ptrs  one process, several devices: buffers[n_ranks] device pointers, devices[n_ranks] their CUDA devices
 *                             (peer access is enabled by the call)
 *   cl_exchange_status        time-outs so far (a peer that does not deliver within ~2 s is counted and read as 0 instead of hanging
 *                             the GPU; synchronises the device) and the number of steps exchanged
 * Not available for building-tiled (wide) districts, central-agent rewards, cl_advance_device, or launches of more than one wave.
 */
int cl_exchange_create(cl_env* env, int32_t n_ranks, int32_t rank, void* ipc_handle_out, void** buffer_out);
int cl_exchange_connect(cl_env* env, const void* ipc_handles);
int cl_exchange_connect_ptrs(cl_env* env, void* const* buffers, const int32_t* devices);
int cl_exchange_status(cl_env* env, int32_t* timeouts, uint32_t* epoch);

/*
 * The observation rows shared by every env after a step: with stale_observations (reference parity) and one
 * episode window for all envs the E rows of `obs` are identical, so a caller that moves observations across PCIe can pass
 * obs = NULL to cl_step / cl_rollout and fetch ONE row per time step instead (CityLearnEnv.observations,
 * citylearn/citylearn.py:451-485).
 *   rows  dev [n_rows][L]   observations at time steps first_time_step .. first_time_step + n_rows - 1, all within [1, T - 1]
 */
int cl_obs_rows(cl_env* env, int32_t first_time_step, int32_t n_rows, float* rows, cl_stream stream);

/* Current time step t of the handle (host value; steps done since the last reset). */
int cl_time_step(const cl_env* env, int32_t* t);

/* Mutable state as an opaque float blob (checkpoint / resume).  cl_state_size returns the number of bytes. */
int cl_state_size(const cl_env* env, size_t* bytes);
int cl_get_state(cl_env* env, void* dst_dev, cl_stream stream);
int cl_set_state(cl_env* env, const void* src_dev, int32_t time_step, cl_stream stream);

/*
 * One step end to end from HOST actions in one call (CityLearnEnv.step as a host-side training loop sees it, citylearn.py:978-1056):
 * copies actions_host [E][action_dim] to actions_dev, runs cl_step into obs_dev (NULL: not materialised) / reward_dev / district_dev
 * (NULL allowed), writes the observation row every env shares at the new time step to row_dev [L] (NULL: skip; needs reference-parity
 * observations and one episode window), copies d2h_bytes from d2h_src_dev to d2h_dst_host (the caller lays its result buffers out so
 * that ONE range covers what it wants back, e.g. [reward | row]) and synchronises the stream.  With pinned host memory both copies are
 * asynchronous DMA transfers; pageable memory works but is staged by the driver.
 * in_place: bit flags, each honoured only when possible (else the copy / synchronisation is used):
 *   1  the kernel reads the actions from actions_host directly (page-locked, device-addressable memory: cudaHostAlloc / cudaHostRegister)
 *   2  the kernels write the results to d2h_dst_host directly (page-locked; obs_dev NULL; the result range holds exactly reward_dev
 *      (+ row_dev)) - posted PCIe stores inside the step instead of a DMA copy after it; the device buffers of the range are NOT written
 *   4  completion is a sequence number a final one-thread kernel stores into page-locked memory, polled by the host, instead of
 *      cudaStreamSynchronize's wake-up
 */
int cl_step_host(cl_env* env, const float* actions_host, float* actions_dev, float* obs_dev, float* reward_dev, float* district_dev,
                 float* row_dev, const void* d2h_src_dev, void* d2h_dst_host, size_t d2h_bytes, int32_t in_place, cl_stream stream);

/* Per-vehicle SOC entries of every env: soc_prev, soc dev [E][n_ev] float (soc[t-1], soc[t]) - diagnostics / parity tests. */
int cl_ev_read(cl_env* env, float* soc_prev_dev, float* soc_dev, cl_stream stream);

/* Number of kernels this library has launched on behalf of `env` since creation (bench.py's gpu_launches). */
int cl_launch_count(const cl_env* env, int64_t* n);

/*
 * Wrapper semantics fused into the kernels (citylearn/wrappers.py:15-238).  Both arguments are HOST arrays
 * copied by the call; NULL restores the identity.  Call between episodes (the call synchronises the device).
 *   obs_transform [L]: per observation column  x -> clip(fn(x) * scale + offset, lo, hi)  with fn = identity, sin(w x) or
 *       cos(w x): NormalizedObservationWrapper (periodic sin / cos columns are separate entries of obs_desc naming the same
 *       source) and ClippedObservationWrapper.
 *   action_range, action_low [action_dim]: the caller's actions are fractions in [0, 1]:  a -> a * range + low
 *       (NormalizedActionWrapper :208-222).
 */
enum cl_obs_fn { CL_OBS_FN_IDENTITY = 0, CL_OBS_FN_SIN = 1, CL_OBS_FN_COS = 2 };
typedef struct cl_obs_transform {
    int32_t fn;        /* cl_obs_fn */
    float w;           /* angular factor 2 pi / x_max of the periodic functions */
    float scale, offset;
    float lo, hi;      /* clip bounds (-inf / +inf: none) */
} cl_obs_transform;
int cl_set_transforms(cl_env* env, const cl_obs_transform* obs_transform, const float* action_range, const float* action_low);

/*
 * Online KPI accumulators for batched envs (citylearn/citylearn.py:1136-1323, cost_function.py:10-388): what
 * CityLearnEnv.evaluate() needs of the action-dependent series, kept per env on the device instead of a per-step history.
 * Control series = the simulated values; baseline = `_without_storage` (net minus the storage devices' consumption; districts with
 * LSTM dynamics are not supported here).  cl_kpi_accumulate is called after every cl_step with that step's `trace` and `district`
 * outputs; cl_reset zeroes the accumulators.  Layouts (doubles):
 * Plain districts (whole envs per thread block, no LSTM dynamics) keep the accumulators INSIDE the step kernel: every cl_step /
 * cl_rollout / cl_advance_device launch updates them (shared memory during the launch, folded into the arrays at its end), no
 * trace is needed and cl_kpi_accumulate is a no-op (cl_kpi_fused reports 1).  Building-tiled districts use cl_kpi_accumulate.
 *   unit [E][B][CL_NKPI_UNIT]: sum max(net,0), sum net, sum max(emission,0), sum max(cost,0) for control, then for the baseline
 *   env  [E][2][CL_NKPI_ENV] : per series (0 control district net, 1 baseline district net) the running ramping / load-factor /
 *                              peak window state of cl_kpi_env
 */
enum cl_kpi_unit { CL_KPI_EC = 0, CL_KPI_ZNE, CL_KPI_EMISSION, CL_KPI_COST, CL_KPI_B_EC, CL_KPI_B_ZNE, CL_KPI_B_EMISSION, CL_KPI_B_COST, CL_NKPI_UNIT };
enum cl_kpi_env {
    CL_KE_N = 0, CL_KE_PREV, CL_KE_RAMP, CL_KE_ALL_MAX,
    CL_KE_D_SUM, CL_KE_D_MAX, CL_KE_D_CNT, CL_KE_D_FIN_LF, CL_KE_D_FIN_PEAK, CL_KE_D_FIN_N,     /* 24-step windows */
    CL_KE_M_SUM, CL_KE_M_MAX, CL_KE_M_CNT, CL_KE_M_FIN_LF, CL_KE_M_FIN_N,                       /* 730-step windows */
    CL_NKPI_ENV
};
int cl_kpi_enable(cl_env* env, int32_t enable);
int cl_kpi_accumulate(cl_env* env, const float* trace, const float* district, cl_stream stream);
int cl_kpi_fused(const cl_env* env, int32_t* fused);
int cl_kpi_read(cl_env* env, double* unit_dev, double* env_dev, cl_stream stream);

/* Launch geometry chosen at cl_create: CTAs per launch, threads per CTA (incl. the helper warp) and building tiles per env
 * (1: a block owns whole envs; > 1: one thread-block cluster per env, one CTA per tile of buildings). */
int cl_launch_geometry(const cl_env* env, int32_t* blocks, int32_t* threads, int32_t* tiles);
/* Resident CTAs per SM of that launch (CUDA occupancy calculator) and its dynamic shared memory per CTA - diagnostics. */
int cl_launch_occupancy(const cl_env* env, int32_t* blocks_per_sm, int32_t* smem_bytes_per_block);

/* Measured FP32 FMA throughput of the current device in TFLOP/s (a microbenchmark of independent FFMA chains): the roofline
 * denominator of the LSTM-dynamics path, which is FP32-FMA bound.  Synchronises the device. */
int cl_measure_fma_peak(double* tflops);

const char* cl_last_error(void);
int cl_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CITYLEARN_B200_H */
