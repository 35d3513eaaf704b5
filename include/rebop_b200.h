/*
 * rebop_b200.h -- C ABI of the B200-native Gillespie (direct method) ensemble engine.
 *
 * This is the drop-in boundary for rebop's hot path.  The reference
 * (Armavica/rebop v0.9.7) has no FFI for this path: it is one Rust crate whose
 * only ABI is the pyo3 module.  Each entry point below therefore cites the Rust
 * item it stands in for (paths relative to the reference tree); a Rust `-sys`
 * crate, the pyo3 module or any other host binds exactly these symbols
 * (INTEGRATION.md shows the bindings).
 *
 * Conventions
 *   - Every function returns a rebop_status (0 = ok).  On failure a thread-local
 *     message is available from rebop_b200_last_error().
 *   - Handles are opaque, created/destroyed by the library, and not internally
 *     synchronised (one thread per handle, like `&mut self`).
 *   - Caller-owned input arrays are copied before the call returns.
 *   - There is no CPU fallback: without a CUDA device every compute entry fails
 *     with REBOP_ERR_CUDA.
 *   - Species counts are carried as int32 on the device (the reference uses
 *     isize); inputs outside the int32 range are rejected.
 */
#ifndef REBOP_B200_H
#define REBOP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  REBOP_OK = 0,
  REBOP_ERR_INVALID = 1,       /* bad argument / shape (the reference's assert! panics, src/gillespie.rs:227-237) */
  REBOP_ERR_OUT_OF_RANGE = 2,  /* species index out of range (src/gillespie.rs:229-233, test issue85_oob) */
  REBOP_ERR_PARSE = 3,         /* "Rate expression not understood" (src/pyo3_gillespie.rs:95-98) */
  REBOP_ERR_MISSING_PARAM = 4, /* "Parameter {s} should have a value" (src/expr.rs:70-72) */
  REBOP_ERR_CUDA = 5,          /* CUDA runtime error, or no device */
  REBOP_ERR_NVRTC = 6,         /* run-time specialisation failed; message holds the compile log */
  REBOP_ERR_LIMIT = 7,         /* network too large for the selected kernel */
  REBOP_ERR_ITER_CAP = 8,      /* a trajectory hit the per-launch iteration cap: state, streams and samples so far are
                                  kept; repeating the call with the same arguments continues it exactly */
  REBOP_ERR_NCCL = 9           /* NCCL missing or failing (rebop_ensemble_stats over several devices) */
} rebop_status;

/* Arithmetic flavour: which of the reference's two engines is reproduced bit for bit. */
typedef enum {
  REBOP_ARITH_API = 0,   /* src/gillespie.rs: factor-by-factor f64 LMA product, select = count(cum < chosen) */
  REBOP_ARITH_MACRO = 1  /* src/gillespie_macro.rs: integer falling factorial, first-match select */
} rebop_arith;

/* Which kernel advances the ensemble. */
typedef enum {
  REBOP_KERNEL_AUTO = 0,   /* build-time kernel if one matches, else NVRTC-specialised, else table-driven */
  REBOP_KERNEL_TABLE = 1,  /* K1: generic kernel, network tables in global memory (one image per batch) */
  REBOP_KERNEL_NVRTC = 2,  /* K2: network-specialised source compiled at run time */
  REBOP_KERNEL_PREBUILT = 3, /* K2: network-specialised source compiled at build time (rebop_sysgen + nvcc) */
  REBOP_KERNEL_PDM = 4     /* K7: partial-propensity kernel for large mass-action networks: the sum of the propensities
                            * factored by owner species (one fused multiply-add per species and reactant pair instead of
                            * two to three multiplies and an add per reaction).  OPT-IN and never picked by AUTO: it is the
                            * direct method with the reference's random-number consumption, but its floating-point sums are
                            * ordered and fused differently from the reference's running sum, so results are statistically
                            * exact, not bit-identical (parity by ensemble statistics, tests/test_pdm.py).  Fills the
                            * "heuristics for large systems" the reference's binding promises and does not have
                            * (python/rebop/gillespie.py:123-126, src/pyo3_gillespie.rs:161).  Needs elementary mass
                            * action (total order <= 2), k >= 0, counts that cannot go negative; no event-log mode. */
} rebop_kernel_kind;

/* Sample type of a batch's time-grid results (the value is the size in bytes).  The reference returns isize
 * (src/pyo3_gillespie.rs:224-237); counts on the device are int32, so REBOP_SAMPLES_I64 only widens.  REBOP_SAMPLES_I16
 * halves the HBM and PCIe traffic of sample-dense runs; a count outside the int16 range makes the run fail with
 * REBOP_ERR_LIMIT instead of wrapping. */
typedef enum { REBOP_SAMPLES_I16 = 2, REBOP_SAMPLES_I32 = 4, REBOP_SAMPLES_I64 = 8 } rebop_sample_dtype;

/* Expression byte-code = post-order walk of `Expr` (src/expr.rs:9-21). */
typedef enum {
  REBOP_OP_CONST = 0, REBOP_OP_SPECIES = 1, REBOP_OP_NEG = 2, REBOP_OP_ADD = 3, REBOP_OP_SUB = 4,
  REBOP_OP_MUL = 5, REBOP_OP_DIV = 6, REBOP_OP_POW = 7, REBOP_OP_MAX = 8, REBOP_OP_MIN = 9,
  REBOP_OP_EXP = 10
} rebop_opcode;

typedef struct {
  int32_t op;    /* rebop_opcode */
  int32_t index; /* species index for REBOP_OP_SPECIES */
  double value;  /* constant for REBOP_OP_CONST */
} rebop_expr_op;

typedef struct rebop_network rebop_network;
typedef struct rebop_batch rebop_batch;
typedef struct rebop_pexpr rebop_pexpr;
typedef struct rebop_system rebop_system;

const char* rebop_b200_last_error(void);
const char* rebop_b200_version(void);
/* Number of CUDA devices visible (0 without a driver); never fails. */
int rebop_b200_device_count(void);

/* ---- network: gillespie::Gillespie::{new, add_reaction}, Rate::{lma, lma_sparse, expr} ---- */

/* Gillespie::new (src/gillespie.rs:168-176): a network over n_species species. */
int rebop_network_create(uint32_t n_species, int arith, rebop_network** out);
void rebop_network_destroy(rebop_network* net);
/* add_reaction(Rate::lma(k, exponents), differences) (src/gillespie.rs:24-26,225-244).
 * exponents, differences: [n_species]. */
int rebop_network_add_reaction_lma(rebop_network* net, double k, const uint32_t* exponents,
                                   const int64_t* differences);
/* add_reaction(Rate::lma_sparse(k, [(index, exponent)...]), differences) (src/gillespie.rs:29-31).
 * Terms are multiplied in the order given, as the reference does. */
int rebop_network_add_reaction_lma_sparse(rebop_network* net, double k, const uint32_t* index,
                                          const uint32_t* exponent, size_t n_terms,
                                          const int64_t* differences);
/* add_reaction(Rate::expr(e), differences) (src/gillespie.rs:33-35). */
int rebop_network_add_reaction_expr(rebop_network* net, const rebop_expr_op* program, size_t n_ops,
                                    const int64_t* differences);
int rebop_network_nb_species(const rebop_network* net, uint32_t* out);   /* src/gillespie.rs:200 */
int rebop_network_nb_reactions(const rebop_network* net, uint32_t* out); /* src/gillespie.rs:210 */

/* Source text of the network-specialised kernel (K2) the engine would compile for this network,
 * and the sm_100a cubin NVRTC produces from it.  Neither needs a GPU.  *needed receives the size
 * in bytes (the source includes its NUL terminator). */
int rebop_network_codegen(const rebop_network* net, char* buf, size_t cap, size_t* needed);
int rebop_network_jit_cubin(const rebop_network* net, char* buf, size_t cap, size_t* needed);
/* The same two for the partial-propensity kernel (REBOP_KERNEL_PDM, see rebop_batch_set_kernel): REBOP_ERR_LIMIT with
 * the reason when the network is not elementary mass action. */
int rebop_network_codegen_pdm(const rebop_network* net, char* buf, size_t cap, size_t* needed);
int rebop_network_jit_cubin_pdm(const rebop_network* net, char* buf, size_t cap, size_t* needed);

/* ---- define_system! (src/gillespie_macro.rs:49-129) ---- */

/* Parses the macro's DSL text: `params...; Name { species, ... } rname: lhs => rhs @ rate ...`.
 * Rate expressions may use parameters, literals, + - * / and parentheses (a rate that names a
 * species -- the macro reads a stale per-call snapshot, src/gillespie_macro.rs:101-104 -- is
 * rejected).  REBOP_ERR_PARSE on failure. */
int rebop_system_parse(const char* dsl_text, rebop_system** out);
void rebop_system_destroy(rebop_system* sys);
int rebop_system_name(const rebop_system* sys, char* buf, size_t cap, size_t* needed);
int rebop_system_counts(const rebop_system* sys, uint32_t* n_params, uint32_t* n_species, uint32_t* n_reactions);
int rebop_system_param_name(const rebop_system* sys, uint32_t i, char* buf, size_t cap, size_t* needed);
int rebop_system_species_name(const rebop_system* sys, uint32_t i, char* buf, size_t cap, size_t* needed);
int rebop_system_reaction_name(const rebop_system* sys, uint32_t i, char* buf, size_t cap, size_t* needed);
/* Name::with_parameters(p...) (src/gillespie_macro.rs:86-95): the network in define_system!
 * arithmetic (REBOP_ARITH_MACRO), reactant factors in the order written (:106). */
int rebop_system_network(const rebop_system* sys, const double* params, size_t n_params, rebop_network** out);
/* The value of every reaction's rate expression for these parameter values (rates: [n_reactions]). */
int rebop_system_rates(const rebop_system* sys, const double* params, size_t n_params, double* rates);
/* Kernels rebop_sysgen + nvcc compiled into this library at build time (the .rsys files under rebop_b200/systems),
 * and whether `net` would run on one of them (same structure; rate constants are launch parameters). */
int rebop_b200_prebuilt_count(void);
int rebop_b200_prebuilt_name(int i, char* buf, size_t cap, size_t* needed);
int rebop_network_has_prebuilt(const rebop_network* net, int* yes);

/* ---- rate expressions: the `PExpr` front end (src/expr.rs:43-273) ---- */

/* "...".parse::<PExpr>() (src/expr.rs:134-141). REBOP_ERR_PARSE on failure. */
int rebop_pexpr_parse(const char* text, rebop_pexpr** out);
void rebop_pexpr_destroy(rebop_pexpr* e);
/* Display for PExpr (src/expr.rs:112-128). Writes a NUL-terminated string; *needed (optional)
 * receives the length including the terminator. */
int rebop_pexpr_format(const rebop_pexpr* e, char* buf, size_t cap, size_t* needed);
/* PExpr::to_expr (src/expr.rs:58-109): species names resolve first, then parameters;
 * REBOP_ERR_MISSING_PARAM ("Parameter {s} should have a value") otherwise.
 * Emits the post-order program; *n_ops receives its length (call with cap = 0 to size it). */
int rebop_pexpr_lower(const rebop_pexpr* e, const char* const* species_names, size_t n_species,
                      const char* const* param_names, const double* param_values, size_t n_params,
                      rebop_expr_op* program, size_t cap, size_t* n_ops);

/* ---- ensemble of trajectories ---- */

/* N instances of Gillespie::new_with_seed(x0, _, seed_n) (src/gillespie.rs:179-187) on `device`.
 * x0: [n_species] when x0_per_trajectory == 0, else [n_traj][n_species].
 * seeds: [n_traj], or NULL for seed_n = seed_base + n. */
int rebop_batch_create(const rebop_network* net, int device, size_t n_traj, const int64_t* x0,
                       int x0_per_trajectory, const uint64_t* seeds, uint64_t seed_base,
                       rebop_batch** out);
void rebop_batch_destroy(rebop_batch* b);
int rebop_batch_set_kernel(rebop_batch* b, int kind);                 /* rebop_kernel_kind */
int rebop_batch_get_kernel(const rebop_batch* b, int* kind);          /* the kernel the last launch used */
/* New rate constants for the mass-action reactions of the batch's network (k: [n_reactions]; entries of
 * expression reactions are ignored).  The struct define_system! generates exposes its parameters as
 * plain fields that may change between advance_until calls (src/gillespie_macro.rs:62-67); trajectories,
 * times and random streams are kept. */
int rebop_batch_set_rates(rebop_batch* b, const double* k, size_t n_reactions);
/* Watchdog: passes of the direct-method loop a trajectory may use per launch (0 = no limit).  A launch that hits it
 * returns REBOP_ERR_ITER_CAP with everything kept on the device; repeating the same call (advance_until with the same
 * tmax, run_grid with the same grid and saved species) continues every trajectory exactly where it stopped --
 * finished trajectories are left alone, so the random streams stay those of an uninterrupted run. */
int rebop_batch_set_max_iters(rebop_batch* b, uint32_t max_iters);
/* How trajectories are mapped to SIMT lanes.
 *   1 = static: thread n runs trajectory n, samples are staged in shared memory and written as full 128-byte lines.
 *   2, 3 = lanes claim trajectories: a resident grid, a lane whose trajectory is finished (or absorbed) claims the
 *       next one from a counter, so no lane idles behind the slowest trajectory of its warp; every trajectory appends
 *       its samples to a record of its own and a second kernel transposes the records into [step][row][trajectory]
 *       (TMA bulk stores), converts them to the batch's sample type and leaves the row sums behind.
 *       2 draws both random words of a pass ahead of the propensities (few samples per event), 3 draws the uniform
 *       once the event is known to fire (many samples per event).
 *   0 = auto (2 or 3 by an estimate of the events per sample).
 * Results are identical: every trajectory owns its state and random stream. */
int rebop_batch_set_schedule(rebop_batch* b, int schedule);
int rebop_batch_get_schedule(const rebop_batch* b, int* schedule_used); /* of the last launch: 1, 2 or 3 */
/* Sample type of run_grid results from now on (rebop_sample_dtype; default REBOP_SAMPLES_I32). */
int rebop_batch_set_sample_dtype(rebop_batch* b, int dtype);
int rebop_batch_get_sample_dtype(const rebop_batch* b, int* dtype);
/* Gillespie::seed (src/gillespie.rs:189-191) for every trajectory. */
int rebop_batch_seed(rebop_batch* b, const uint64_t* seeds, uint64_t seed_base);
/* get/set_time, get/set_species (src/gillespie.rs:246-267). species: [n_traj][n_species]. */
int rebop_batch_get_time(rebop_batch* b, double* t);
int rebop_batch_set_time(rebop_batch* b, double t);
int rebop_batch_get_species(rebop_batch* b, int64_t* species);
int rebop_batch_set_species(rebop_batch* b, const int64_t* species, int per_trajectory);
/* Gillespie::advance_until (src/gillespie.rs:315-344) on every trajectory. */
int rebop_batch_advance_until(rebop_batch* b, double tmax);
/* Gillespie::advance_one_reaction (src/gillespie.rs:270-297) on every trajectory: exactly one pass of the direct
 * method whatever the current time; a trajectory in an absorbing state gets t = +inf (:281-284). */
int rebop_batch_advance_one_reaction(rebop_batch* b);
/* The pyo3 grid loop (src/pyo3_gillespie.rs:197-208): for i in 0..=nb_steps
 * { advance_until(tmax*i/nb_steps); record species[save_idx] }.  save_idx: strictly increasing
 * species indices, or NULL for all species.
 * Samples stay on the device as int32 [nb_steps+1][n_save][ld] (ld >= n_traj).
 * host_out (optional): caller buffer of (nb_steps+1)*n_save*n_traj int32 that receives them
 * densely as [step][save][trajectory]; a large result is then produced in a few segments of consecutive
 * grid points, each copied to the host while the next is simulated (page-locked memory makes that overlap). */
int rebop_batch_run_grid(rebop_batch* b, double tmax, uint32_t nb_steps, const uint32_t* save_idx,
                         uint32_t n_save, int32_t* host_out);
/* Same for a batch of any sample type: host_out (optional) holds (nb_steps+1)*n_save*n_traj samples of the batch's
 * sample type (rebop_batch_set_sample_dtype). */
int rebop_batch_run_grid_typed(rebop_batch* b, double tmax, uint32_t nb_steps, const uint32_t* save_idx,
                               uint32_t n_save, void* host_out);
/* Same into a wider host array: row r of this batch goes to host_out[r * ld .. r * ld + n_traj) (elements of the
 * batch's sample type), so the shards of an ensemble land side by side in one [step][save][all trajectories] array. */
int rebop_batch_run_grid_strided(rebop_batch* b, double tmax, uint32_t nb_steps, const uint32_t* save_idx,
                                 uint32_t n_save, void* host_out, size_t ld);
/* The nb_steps = 0 path of the binding (src/pyo3_gillespie.rs:209-223) for every trajectory:
 * record; while t < tmax { _advance_one_reaction (src/gillespie.rs:275-297); record } -- one row per applied
 * reaction, the last one at or beyond tmax (t = +inf when the state became absorbing).  The log stays on the
 * device: times f64 [total_rows], samples int32 [n_save][total_rows]; trajectory n owns rows
 * offsets[n] .. offsets[n+1].  Needs fewer than 2^32 rows in total. */
int rebop_batch_run_events(rebop_batch* b, double tmax, const uint32_t* save_idx, uint32_t n_save);
int rebop_batch_events_log_size(const rebop_batch* b, uint64_t* total_rows, uint32_t* n_save);
/* offsets: [n_traj + 1], times: [total_rows], samples: [n_save][total_rows]; any of them may be NULL. */
int rebop_batch_events_log_host(rebop_batch* b, uint64_t* offsets, double* times, int32_t* samples);
/* Device view of the last run_grid's samples: [n_rows][ld] elements of the batch's sample type. */
int rebop_batch_samples_device(const rebop_batch* b, const void** dev_ptr, size_t* ld, uint32_t* n_rows);
/* Copy the last run_grid's samples to the host as [step][save][trajectory]: in the batch's own sample type, or as
 * int32 / int64 whatever that type is (converted on the device, a chunk of rows at a time). */
int rebop_batch_samples_host(rebop_batch* b, void* out);
int rebop_batch_samples_host_i32(rebop_batch* b, int32_t* out);
int rebop_batch_samples_host_i64(rebop_batch* b, int64_t* out);
/* Same as _i32 into a wider host array: row r of this batch goes to out[r * ld .. r * ld + n_traj), so the
 * shards of an ensemble (one batch per GPU) land side by side in one [step][save][all trajectories] array
 * with no intermediate copy; out points at the first trajectory of this batch in row 0. */
int rebop_batch_samples_host_i32_strided(rebop_batch* b, int32_t* out, size_t ld);
int rebop_batch_samples_host_strided(rebop_batch* b, void* out, size_t ld); /* in the batch's own sample type */
/* K4: per (step, saved species) sum and sum of squares over the trajectories of this batch,
 * as exact integers (so that sums over GPUs are order-independent). sum, sumsq: [n_rows]. */
int rebop_batch_sample_sums(rebop_batch* b, int64_t* sum, uint64_t* sumsq);
/* Same, left on the device for a following all-reduce: int64 [2][n_rows] (sums, then squares). */
int rebop_batch_sample_sums_device(rebop_batch* b, const int64_t** dev_ptr, uint32_t* n_rows);
/* Applied reactions since creation, and during the last launch. */
int rebop_batch_events(rebop_batch* b, uint64_t* total, uint64_t* last_launch);
/* Lane slots of the last launch: 32 x the loop iterations each warp executed.  events / lane slots is
 * the fraction of SIMT lanes that applied a reaction (the rest idled on finished trajectories,
 * rejected ziggurat draws or grid crossings). */
int rebop_batch_lane_slots(rebop_batch* b, uint64_t* last_launch);
/* Device time of the ensemble loop of the last advance_until / run_grid call in milliseconds (CUDA events on the
 * batch's stream), and of the kernels that brought its samples into the result layout. */
int rebop_batch_last_kernel_ms(rebop_batch* b, float* ms);
int rebop_batch_last_finish_ms(rebop_batch* b, float* ms);
int rebop_batch_size(const rebop_batch* b, size_t* n_traj);
/* Block until the device has finished the batch's queued work.  On a caller-owned stream (rebop_batch_set_stream)
 * advance_until / run_grid / advance_one_reaction do not block the host; this call then reports what they would
 * have (REBOP_ERR_ITER_CAP, REBOP_ERR_LIMIT) and accounts their events. */
int rebop_batch_synchronize(rebop_batch* b);
/* The CUDA stream (cudaStream_t / CUstream) every launch and copy of this batch is issued on, so
 * that a host framework can order its own work (events, collectives) against it. */
int rebop_batch_get_stream(const rebop_batch* b, void** stream);
/* Issue this batch's work on a caller-owned stream instead (NULL restores the batch's own). */
int rebop_batch_set_stream(rebop_batch* b, void* stream);

/* ---- one ensemble over several GPUs of one node ---- */

/* N trajectories sharded over `devices` in contiguous ranges [g*N/G, (g+1)*N/G): one batch per device, one host
 * worker thread per device for the duration of a call, no data-path collective (trajectories are independent; the
 * reference's only ensemble is the host loop of examples/sir.rs:11-22).  Every trajectory carries its own seed
 * (seeds[n], or seed_base + n), so results do not depend on the number of devices.  Arguments as for
 * rebop_batch_create, indexed over the whole ensemble. */
typedef struct rebop_ensemble rebop_ensemble;
int rebop_ensemble_create(const rebop_network* net, const int* devices, int n_devices, size_t n_traj, const int64_t* x0,
                          int x0_per_trajectory, const uint64_t* seeds, uint64_t seed_base, rebop_ensemble** out);
void rebop_ensemble_destroy(rebop_ensemble* e);
int rebop_ensemble_size(const rebop_ensemble* e, size_t* n_traj);
int rebop_ensemble_shards(const rebop_ensemble* e, int* n_shards);
/* The batch of shard g (owned by the ensemble; NULL if its range is empty), its device and trajectory range. */
int rebop_ensemble_shard(const rebop_ensemble* e, int g, rebop_batch** batch, int* device, size_t* first, size_t* count);
/* The batch setters and steppers of the same name, applied to every shard concurrently. */
int rebop_ensemble_set_kernel(rebop_ensemble* e, int kind);
int rebop_ensemble_set_schedule(rebop_ensemble* e, int schedule);
int rebop_ensemble_set_sample_dtype(rebop_ensemble* e, int dtype);
int rebop_ensemble_set_max_iters(rebop_ensemble* e, uint32_t max_iters);
int rebop_ensemble_set_rates(rebop_ensemble* e, const double* k, size_t n_reactions);
int rebop_ensemble_set_time(rebop_ensemble* e, double t);
int rebop_ensemble_set_species(rebop_ensemble* e, const int64_t* species, int per_trajectory);
int rebop_ensemble_seed(rebop_ensemble* e, const uint64_t* seeds, uint64_t seed_base);
int rebop_ensemble_advance_until(rebop_ensemble* e, double tmax);
int rebop_ensemble_advance_one_reaction(rebop_ensemble* e);
/* rebop_batch_run_grid on every shard; host_out (optional): [step][save][all n_traj trajectories] in the ensemble's
 * sample type, every shard writing its own columns. */
int rebop_ensemble_run_grid(rebop_ensemble* e, double tmax, uint32_t nb_steps, const uint32_t* save_idx, uint32_t n_save,
                            void* host_out);
int rebop_ensemble_samples_host(rebop_ensemble* e, void* out);
/* Ensemble statistics of the last run_grid per (step, saved species): every device reduces its shard to exact int64
 * sums and sums of squares, ncclAllReduce(ncclInt64, ncclSum) over an ncclCommInitAll communicator combines them (no
 * NCCL call with one device), and a device kernel turns them into mean and unbiased variance with 128-bit integer
 * arithmetic.  Integer sums make the result independent of the device count and of the reduction order.
 * NCCL is opened with dlopen (REBOP_B200_NCCL_LIB overrides the path); REBOP_ERR_NCCL if it is missing. */
int rebop_ensemble_stats(rebop_ensemble* e, double* mean, double* var);
int rebop_ensemble_sums(rebop_ensemble* e, int64_t* sum, uint64_t* sumsq);
int rebop_ensemble_events(rebop_ensemble* e, uint64_t* total, uint64_t* last_call);
int rebop_ensemble_last_kernel_ms(rebop_ensemble* e, float* max_ms);

/* ---- host memory ---- */

/* Page-locked host memory for result buffers (device-to-host copies into it run at full PCIe
 * rate and asynchronously).  Stands in for the Vec the reference returns
 * (src/pyo3_gillespie.rs:224-237); the caller frees it with rebop_b200_host_free. */
int rebop_b200_host_alloc(size_t bytes, void** out);
int rebop_b200_host_free(void* p);

/* ---- measurement helpers ---- */

/* Number of kernels this library has launched in the calling process so far (all batches). */
uint64_t rebop_b200_kernel_launches(void);

/* Sustained non-fused FP64 issue rate of `device` in operations per second (independent
 * DADD/DMUL chains, all SMs) -- the denominator of the per-event FP64 roofline. */
int rebop_b200_measure_fp64_rate(int device, double* ops_per_second, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* REBOP_B200_H */
