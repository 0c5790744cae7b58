/* rydcz.h -- C ABI of the B200 Rydberg-CZ gate evaluator.
 *
 * The reference has no FFI: its boundary for this path is the Python function
 *   simulate_CZ_gate(...)            RG/simulation.py:2534-3676
 * (RG = src/qpu_simulator/micro_physics/neutral_atoms/rydberg_gates).  Inside it, the part this
 * library replaces is everything from operator assembly to the fidelity numbers:
 *   build_full_hamiltonian / build_phase_modulated_hamiltonian   RG/hamiltonians.py:1057, 1179
 *   build_all_noise_operators                                    RG/noise_models.py:1449
 *   evolve_two_pulse_lp / evolve_bangbang_jp /
 *   evolve_smooth_sinusoidal_jp / evolve_shaped_pulse            RG/simulation.py:740, 1795, 1502, 2099
 *     (each a Python loop of QuTiP mesolve calls,                RG/simulation.py:689)
 *   compute_CZ_fidelity                                          RG/simulation.py:225
 * A maintainer binds these entry points with ctypes.
 *
 * All arrays are IEEE double unless stated, row-major, caller-owned.  One "point" is one gate
 * evaluation (all four computational-basis inputs through the whole pulse).
 *
 *   points  [n_points][RYDCZ_POINT_STRIDE]   per-point constants, see enum below
 *   segs    [n_points][max_seg][4]           piecewise-constant pulse: dt (s), Re Omega, Im Omega,
 *                                            Delta (rad/s).  Built by the caller exactly as the
 *                                            reference's evolve_* loops build their Hamiltonians.
 *   n_seg   [n_points] (int32)               segments actually used by each point (ragged batches)
 *   out     [n_points][RYDCZ_OUT_STRIDE]     results, see enum below
 *   states  [n_points][rydcz_state_stride(dim)] optional raw reduced states (rho and coherences);
 *                                            NULL to skip.  Layout: rydcz_state_offset().
 *
 * Every function returns 0 on success, non-zero on failure; rydcz_last_error() gives the text.
 * Per-point numerical trouble never raises: it shows up as NaN in that point's row.
 */
#ifndef RYDCZ_H
#define RYDCZ_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RYDCZ_POINT_STRIDE 12
#define RYDCZ_SEG_STRIDE 4
#define RYDCZ_OUT_STRIDE 16

enum rydcz_point_slot {
  RYDCZ_PT_V = 0,         /* blockade shift V = C6/R^6 (rad/s)           RG/simulation.py:2958-2959 */
  RYDCZ_PT_DELTA1 = 1,    /* delta_zeeman + delta_stark on |1> (rad/s)    RG/simulation.py:3129-3159 */
  RYDCZ_PT_OFF_A = 2,     /* extra detuning of atom A (rad/s), 0 in the reference                    */
  RYDCZ_PT_OFF_B = 3,     /* extra detuning of atom B                                                */
  RYDCZ_PT_G_R1 = 4,      /* Gamma(r->1) = gamma_r * branching_1          RG/noise_models.py:1237    */
  RYDCZ_PT_G_R0 = 5,      /* Gamma(r->0): all channels ending in |0>      RG/noise_models.py:1242-1251, 1583-1592 */
  RYDCZ_PT_G_PHI = 6,     /* dephasing rate on |r><r|                     RG/noise_models.py:1579    */
  RYDCZ_PT_G_SC = 7,      /* scattering rate on |1><1|                    RG/noise_models.py:1589    */
  RYDCZ_PT_G_MJ = 8,      /* r+ <-> r- mixing rate (dim = 4 only)         RG/noise_models.py:1287    */
  RYDCZ_PT_SIG_DET = 9,   /* Monte-Carlo shots: std of per-atom Doppler detuning k_eff v_rms (rad/s) */
  RYDCZ_PT_SIG_RELR = 10, /* Monte-Carlo shots: std of dR/R0 = sqrt(2) sigma_r / R0                  */
  RYDCZ_PT_RESERVED = 11
};

enum rydcz_out_slot {
  RYDCZ_OUT_AVG_F = 0,        /* avg_fidelity                           RG/simulation.py:631        */
  RYDCZ_OUT_F00 = 1,
  RYDCZ_OUT_F01 = 2,
  RYDCZ_OUT_F10 = 3,
  RYDCZ_OUT_F11 = 4,          /* F11 population x cz_phase_fidelity     RG/simulation.py:619-629    */
  RYDCZ_OUT_F11_POP = 5,
  RYDCZ_OUT_CZ_PHASE_FID = 6, /* cos^2(phase_error/2)                   RG/simulation.py:611        */
  RYDCZ_OUT_CTRL_PHASE = 7,   /* phi11 - phi01 - phi10 + phi00 in [-pi, pi)   RG/simulation.py:607-608 */
  RYDCZ_OUT_PHASE_ERR = 8,    /* min(|c - pi|, |c + pi|)                RG/simulation.py:610        */
  RYDCZ_OUT_PHI01 = 9,
  RYDCZ_OUT_PHI10 = 10,
  RYDCZ_OUT_PHI11 = 11,
  RYDCZ_OUT_LOSS = 12,        /* |11> input: P(an atom ended in |0>) -- where every loss channel lands */
  RYDCZ_OUT_RYD = 13,         /* |11> input: P(an atom is left in a Rydberg level)                     */
  RYDCZ_OUT_STEPS = 14,       /* sparse matvecs performed (work counter)                              */
  RYDCZ_OUT_FLOPS = 15        /* useful FP64 flops performed by this point                            */
};

typedef struct rydcz_ctx rydcz_ctx;

/* per-point status of the last evaluation call (rydcz_point_status): which path evaluated the point */
enum rydcz_status {
  RYDCZ_STATUS_PLAIN = 0,            /* single-interval Chebyshev, general two-atom systems              */
  RYDCZ_STATUS_PLAIN_FOLDED = 1,     /* single-interval Chebyshev, exchange-folded systems               */
  RYDCZ_STATUS_CLUSTER_GENERAL = 2,  /* clustered-spectrum expansion, unequal atoms / Monte-Carlo shot   */
  RYDCZ_STATUS_CLUSTER = 3,          /* clustered-spectrum expansion, identical atoms                    */
  RYDCZ_STATUS_REPEAT = 5,           /* one propagator for a phase-modulated pulse (smooth JP)           */
  RYDCZ_STATUS_UNSUPPORTED = 6,      /* not evaluated: the row is NaN (no case left since round 2)       */
  RYDCZ_STATUS_HANDED_BACK = 0x10    /* flag: the clustered path's invariant check failed and the plain
                                        kernel re-evaluated the point inside the same call             */
};
/* call-level flags of the last evaluation call (rydcz_call_flags) */
#define RYDCZ_FLAG_NO_CLUSTER_SCRATCH 1   /* the coefficient scratch of the clustered path could not be allocated
                                             (or would exceed 48 GB): the whole call ran on the plain path */
#define RYDCZ_FLAG_UNSUPPORTED_POINTS 2   /* some rows are NaN because no kernel evaluates them (see status) */

/* library / layout queries (no GPU needed) */
int rydcz_version(void);
int rydcz_point_stride(void);
int rydcz_seg_stride(void);
int rydcz_out_stride(void);
int rydcz_point_stride_dim(int dim);                    /* 12 (dim 3, 4), 20 (dim 5), -1 otherwise */
int rydcz_state_stride(int dim);                        /* 51 (dim 3), 64 (dim 4), 129 (dim 5), -1 otherwise */
int rydcz_state_offset(int dim, const char* system);    /* "rho11","rho01","rho10","coh11","coh01","coh10" */
int rydcz_system_size(int dim, const char* system);     /* real variables of that reduced system */
int rydcz_system_nnz(int dim, const char* system);      /* FMAs per sparse matvec of that system */

/* context: one per (process, GPU); owns a stream, timing events and scratch buffers */
int rydcz_create(int device, rydcz_ctx** ctx);
void rydcz_destroy(rydcz_ctx* ctx);
const char* rydcz_last_error(const rydcz_ctx* ctx);     /* ctx may be NULL: last create() error */
int rydcz_set_sort(rydcz_ctx* ctx, int enable);         /* order points by estimated work (default 1) */
/* Clustered-spectrum propagation (csrc/rydcz_cluster.cuh) for points with a strong blockade (identical atoms:
 * 3- and 4-level models; unequal atoms and Monte-Carlo shots: 3-level model): 0 = never, 1 = where the planner
 * estimates it cheaper than the plain Chebyshev path (default), 2 = wherever it is eligible.  Same results to
 * rounding; a point whose trace / norm check fails on that path is re-evaluated on the plain path inside the
 * same call.  Device memory: 16.6 KB of coefficient records per point (per shot with shots), kept by the
 * context; if it cannot be allocated (or would exceed 48 GB) the call stays on the plain path. */
int rydcz_set_cluster(rydcz_ctx* ctx, int mode);
/* Shared coefficient table of the clustered path (csrc/rydcz_cluster.cuh): the coefficient records of a (sub)step
 * depend only on (h Vc, rho / Vc^3), and a (sub)step may use the records of the nearest grid entry EXACTLY if its
 * operator is rescaled by lambda = h Vc / theta_n and the entry's clusters are a little wider -- so slots that
 * share the blockade and the drive (every candidate of an optimiser population, every chunk of a sweep at fixed
 * apparatus) share a few thousand table entries instead of generating 1-8 sets of records per point.
 * 0 = off (every slot generates its own records, round-1 behaviour), 1 = table entries generated once per
 * evaluation call (nothing is carried from one call to the next), 2 = entries kept across calls (default: an
 * entry is a pure function of its index; an optimiser's generations and a sweep's calls re-use them).  Device memory: 1.2 GB, allocated at the first clustered call; if
 * that fails the context silently behaves like mode 0 (results identical to rounding in every mode). */
int rydcz_set_table(rydcz_ctx* ctx, int mode);
/* table entries generated / slots of clustered points that took their records from the table, since the last
 * call of this function (synchronises the device; any pointer may be NULL) */
int rydcz_table_stats(rydcz_ctx* ctx, int64_t* entries_generated, int64_t* slots_shared);
/* points planned onto the clustered path / handed back to the plain path / 32-byte coefficient records
 * written for them since the last call of this function (synchronises the device; any pointer may be NULL) */
int rydcz_cluster_stats(rydcz_ctx* ctx, int64_t* planned, int64_t* handed_back, int64_t* records);
/* Points evaluated on the repeated-segment path (csrc/rydcz_repeat.cuh: identical atoms, 3-level model, at
 * least 32 active segments that differ only in the laser phase -- the smooth Jandura-Pupillo protocol) since
 * the last call of this function.  The path is part of the automatic choice: rydcz_set_cluster(0) disables it. */
int rydcz_repeat_stats(rydcz_ctx* ctx, int64_t* points);
/* Status column: the path that evaluated each of the first n items of the last evaluation
 * call (points; shots when a Monte-Carlo call was planned per shot), an enum rydcz_status value, possibly with
 * RYDCZ_STATUS_HANDED_BACK or'ed in.  Nothing degrades silently: a call that could not use the clustered path,
 * or that left rows unevaluated, says so in rydcz_call_flags.  Both synchronise the device. */
int rydcz_point_status(rydcz_ctx* ctx, int64_t n, int32_t* h_status);
int rydcz_call_flags(rydcz_ctx* ctx, int* flags, int64_t* unsupported_points);

/* One batch of gate evaluations, inputs and outputs resident in device memory.
 * dim: 3 (|0>,|1>,|r>) or 4 (|0>,|1>,|r+>,|r->).  noisy: 0 -> Schroedinger path with the
 * pure-state epilogue (QuTiP's c_ops=[] branch); 1 -> Lindblad path.  stream: a cudaStream_t
 * (as void*) to enqueue on, or NULL for the context's own stream.  Asynchronous.
 * Streams: a context owns ONE set of planning scratch (work order, kinds, coefficient records).  The
 * asynchronous entry points (*_device) may be given different streams from call to call: each call waits, on
 * the device, for the previous call of the same context (an event recorded at its end), so calls of one context
 * never overlap.  For concurrent evaluation on several streams create one context per stream.  Host threads
 * must not call into one context concurrently. */
int rydcz_evaluate_device(rydcz_ctx* ctx, int dim, int noisy, int64_t n_points,
                          const double* d_points, const double* d_segs, const int32_t* d_n_seg,
                          int max_seg, double* d_out, double* d_states, void* stream);

/* Same with host buffers: copies in, runs, copies out, and returns when the results are in
 * h_out (and h_states).  This is the call the reference-side binding makes. */
int rydcz_evaluate_host(rydcz_ctx* ctx, int dim, int noisy, int64_t n_points,
                        const double* h_points, const double* h_segs, const int32_t* h_n_seg,
                        int max_seg, double* h_out, double* h_states);

/* Monte-Carlo shot averaging (extension; the reference folds Doppler and position noise into
 * dephasing rates, RG/trap_physics.py:1206-1288, 1064-1115).  Shot s of point i draws, from
 * Philox4x32-10 keyed by (seed; i, s), two Doppler detunings N(0, SIG_DET^2) and one relative
 * displacement eps ~ N(0, SIG_RELR^2) with V -> V / (1 + eps)^6, and is evaluated like any
 * point.  d_out_mean[n_points][16] receives the shot mean of every slot (STEPS and FLOPS: the
 * sum); d_out_shots [n_points][n_shots][16] is optional. */
int rydcz_evaluate_shots_device(rydcz_ctx* ctx, int dim, int noisy, int64_t n_points, int n_shots,
                                uint64_t seed, const double* d_points, const double* d_segs,
                                const int32_t* d_n_seg, int max_seg, double* d_out_mean,
                                double* d_out_shots, void* stream);
/* Sharded Monte-Carlo calls: shot s of point i is keyed by (seed; first_point + i, s).  A rank that evaluates the
 * points [lo, hi) of a larger batch sets first_point = lo and draws exactly the shots the unsharded call draws
 * (default 0; stays set until changed). */
int rydcz_set_point_base(rydcz_ctx* ctx, int64_t first_point);
int rydcz_evaluate_shots_host(rydcz_ctx* ctx, int dim, int noisy, int64_t n_points, int n_shots,
                              uint64_t seed, const double* h_points, const double* h_segs,
                              const int32_t* h_n_seg, int max_seg, double* h_out_mean,
                              double* h_out_shots);

/* ---- simulate: scalar prologue on the device, then evaluate ---------------------------------
 * Replaces, for a batch, steps 1-13 of simulate_CZ_gate (RG/simulation.py:2820-3572): the per-point
 * physics of laser_physics / trap_physics / protocols / noise_models part 1 (Rabi chain, blockade,
 * LP table and phase jump, trap-dependent rates, coherent shifts, pulse schedule) runs in
 * prologue_kernel from RYDCZ_APP_STRIDE apparatus slots per point, stored SLOT-MAJOR:
 * app[RYDCZ_APP_STRIDE][n_points] (contiguous columns; NaN = "None", the reference's default;
 * slot list: enum AppSlot in csrc/rydcz_prologue.cuh); species- and
 * n-level constants arrive once per call in `model` (RYDCZ_MODEL_STRIDE doubles, filled by the
 * Python mirror: rydberg_gates/device_prologue.py).  Protocols: square LP, JP bang-bang, smooth JP.
 * `scalars` [n][RYDCZ_SCALAR_STRIDE] (optional) returns Omega, V, tau_total, tau_single, Delta,
 * Re xi, Im xi, Omega*tau.  n_shots > 0 -> shot means as in rydcz_evaluate_shots_*.            */
#define RYDCZ_APP_STRIDE 40
#define RYDCZ_MODEL_STRIDE 32
#define RYDCZ_SCALAR_STRIDE 8
int rydcz_app_stride(void);
int rydcz_model_stride(void);
int rydcz_scalar_stride(void);
int rydcz_prologue_device(rydcz_ctx* ctx, const double* h_model, int64_t n_points, const double* d_app,
                          double* d_points, double* d_segs, int32_t* d_n_seg, double* d_scalars,
                          void* stream);
int rydcz_simulate_device(rydcz_ctx* ctx, const double* h_model, int64_t n_points, const double* d_app,
                          int n_shots, uint64_t seed, double* d_out, double* d_scalars, double* d_states,
                          void* stream);
int rydcz_simulate_host(rydcz_ctx* ctx, const double* h_model, int64_t n_points, const double* h_app,
                        int n_shots, uint64_t seed, double* h_out, double* h_scalars, double* h_states);

/* Column form of rydcz_simulate_host, pipelined -- the call simulate_CZ_gate_batch makes.
 * Slot k of point i is h_cols[k][i * col_stride[k]] when h_cols[k] != NULL, else the batch-wide
 * constant h_const[k] (h_cols, col_stride, h_const: RYDCZ_APP_STRIDE entries; col_stride NULL = all
 * 1; strides in doubles, so a column of a row-major [n][m] array is {ptr + k, m}).  Only varying
 * slots cross PCIe.  The batch is cut into chunks (rydcz_set_chunk, default 32768 evaluations) and
 * each chunk's packing, H2D copy, kernels and D2H copy run on one of eight streams, so copies and
 * host work overlap the kernels of the neighbouring chunks.  Results are identical to the
 * unchunked call (Monte-Carlo shots are keyed by the point's index in the whole batch).  h_out /
 * h_scalars / h_states should be page-locked for the D2H copies to overlap. */
int rydcz_simulate_host_cols(rydcz_ctx* ctx, const double* h_model, int64_t n_points,
                             const double* const* h_cols, const int64_t* col_stride, const double* h_const,
                             int n_shots, uint64_t seed, double* h_out, double* h_scalars, double* h_states);
int rydcz_set_chunk(rydcz_ctx* ctx, int64_t points_per_chunk);
/* Host threads that gather the varying columns into the page-locked staging (the caller's thread included).
 * Default: (CPUs this process may run on) / LOCAL_WORLD_SIZE - 1, clamped to 1..4 -- one process per GPU must not
 * oversubscribe the host; env RYDCZ_PACK_THREADS overrides.  The workers sleep between calls. */
int rydcz_set_pack_threads(rydcz_ctx* ctx, int threads);
/* Bind the calling thread, and threads created by it afterwards, to the CPUs local to the context's GPU
 * (/sys/bus/pci/devices/<gpu>/local_cpulist): call once per rank, right after rydcz_create, on multi-socket hosts.
 * *n_cpus = CPUs in the new mask, 0 if the topology could not be read (then nothing changes). */
int rydcz_bind_host_to_gpu(rydcz_ctx* ctx, int* n_cpus);
/* How the pipelined host entry points wait for the GPU: 0 = the driver's default (the calling thread spins on a
 * core until the streams drain: lowest latency, one process per host), 1 = sleep on blocking-sync events (the
 * core goes back to the other ranks of a shared host).  Env RYDCZ_HOST_WAIT=spin|sleep sets the initial mode. */
int rydcz_set_host_wait(rydcz_ctx* ctx, int mode);
/* The same call returning only the result columns the caller reads, COLUMN-MAJOR: h_out_cm[j * n_points + i] =
 * out[i][out_slots[j]] (enum rydcz_out_slot), h_scalars_cm[j * n_points + i] = scalars[i][scalar_slots[j]].
 * The device gathers the columns, so D2H traffic is 8 (n_out + n_scalars) bytes per point instead of 192
 * (the reference's callers read ~6 numbers per evaluation, RG/optimize_cz_gate.py:434-451). */
int rydcz_simulate_host_select(rydcz_ctx* ctx, const double* h_model, int64_t n_points,
                               const double* const* h_cols, const int64_t* col_stride, const double* h_const,
                               int n_shots, uint64_t seed, const int32_t* out_slots, int n_out, double* h_out_cm,
                               const int32_t* scalar_slots, int n_scalars, double* h_scalars_cm);

/* page-locked host buffers for the *_host entry points (plain malloc memory works too, slower) */
int rydcz_pinned_alloc(size_t bytes, void** ptr);
int rydcz_pinned_free(void* ptr);

/* measurement */
int rydcz_last_kernel_ms(rydcz_ctx* ctx, float* ms);    /* device time of the last evaluate kernel */
int rydcz_last_coef_ms(rydcz_ctx* ctx, float* ms);      /* ... of the last clustered-path coefficient kernel (0 if none) */
int rydcz_launch_count(const rydcz_ctx* ctx, int64_t* n); /* kernels launched through this context */
/* register-resident FP64 FMA throughput of this GPU (the roofline denominator), TFLOP/s */
int rydcz_fp64_peak(rydcz_ctx* ctx, double* dfma_tflops, double* dmma_tflops);
/* the same pipe driven by the hot loop's own instruction mix (the folded 3-level rho_11 matvec, three
 * distinct register operands per DFMA, two CTAs of 128 threads per SM): the ceiling for a kernel of
 * this shape, TFLOP/s */
int rydcz_fp64_matvec_peak(rydcz_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* RYDCZ_H */
