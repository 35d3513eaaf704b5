// ssa_params.h -- launch parameters of the ensemble kernels (host and device; no std includes,
// so that NVRTC can compile it as is).
#pragma once

typedef unsigned long long rb_u64;
typedef long long rb_i64;
typedef unsigned int rb_u32;

#define RB_FULL_MASK 0xffffffffu
// Static shared memory of the ensemble loop (ziggurat tables: 256 (X[i], X[i+1]) pairs, 256 (F[i], F[i+1]) pairs,
// 256 chord slopes)
#define RB_STATIC_SMEM_BYTES ((512 + 512 + 256) * 8)
// RB_ZIG_WIDE kernels (the register-resident specialised form) hold five 16-byte entries per layer instead
#define RB_ZIG_WIDE_EXTRA_BYTES (256 * 16 * 5 - RB_STATIC_SMEM_BYTES)

// Rate constants carried in the launch parameters (bounds the reactions a specialised kernel can have).
#define RB_MAX_K 1024

// Record of one reaction in SsaRunParams::gtab (8 words, built by the engine at launch):
//   w0,w1  rate constant (f64)            w2  species of term 0 | species of term 1 << 16
//   w3     exponent 0 | exponent 1 << 8 | number of terms << 16 | (more than four species change) << 24
//          number of terms = 0xff: not a record reaction (expression rate or more than two terms)
//   w4,w5  jump species 0..3 (u16 each)   w6,w7  jump differences 0..3 (i16 each, 0 = unused)
// followed, after the last reaction, by the saved-species list (one word each).
#define RB_GTAB_WORDS_PER_REACTION 8

// Status bits written to SsaRunParams::status.
#define RB_STATUS_ITER_CAP 1u  // a trajectory hit max_iters before reaching its last grid point
#define RB_STATUS_NARROW 2u    // a sample did not fit the requested narrow sample type (rb_samples_finish)

// SsaRunParams::progress[traj]: grid points of the current call the trajectory has written itself so far (low 31
// bits) and whether it has reached the call's last grid point (bit 31).  Zeroed by the engine before a fresh
// call; a launch that follows REBOP_ERR_ITER_CAP reads it (SsaRunParams::resuming) so that every trajectory
// continues exactly where it stopped and finished ones are left alone.  A trajectory that reached an absorbing
// state stops writing rows: every later grid point repeats its last row (rb_samples_finish fills them in).
#define RB_PROGRESS_DONE 0x80000000u
#define RB_PROGRESS_ROWS 0x7fffffffu

// Schedules = compile-time variants of the ensemble loop (rb_ssa_loop<Net, MODE>).
#define RB_MODE_STATIC 0  // thread n runs trajectory n; samples ring-staged, written as [step][row][trajectory] lines
#define RB_MODE_SPARSE 1  // lanes claim trajectories; both random words drawn ahead of the propensities (few samples per event)
#define RB_MODE_DENSE 2   // lanes claim trajectories; the uniform is drawn once the event is known to fire (many samples per event)

// Launch parameters (passed by value; lives in the constant bank).
struct SsaRunParams {
  int* x;                // [S][ldn] species counts, trajectory-contiguous
  double* t;             // [ldn] current time of each trajectory
  rb_u64* rng;           // [4][ldn] xoshiro256++ state (seeded by the engine's rb_seed_kernel before the first launch)
  int* out;              // samples, or null.  static: [(step-step_first)][n_save][ldn].  lanes-claim-trajectories
                         // schedules: raw rows [traj][step-step_first][n_save], one contiguous record per trajectory
                         // (rb_samples_finish transposes them into [step][row][trajectory])
  rb_u64* events;        // [0] += applied reactions; [2] += lane slots (32 x loop iterations of each warp)
  rb_u32* status;        // [1] |= RB_STATUS_*
  double tmax;
  const double* grid_t;  // [step_last - step_first + 1] grid times t_i = (tmax * i) / nb_steps, or null: single target tmax
  rb_u32 n_traj, ldn;
  rb_u32 nb_steps;       // 0 => single target tmax (Gillespie::advance_until)
  rb_u32 step_first, step_last;  // grid points handled by this launch (inclusive)
  rb_u32 n_save;         // rows per sample
  rb_u32 ring_depth;     // power of two, 1..32: grid points a warp can stage in shared memory
  rb_u32 max_iters;      // per-trajectory loop-iteration cap for this launch, rounded up to whole ticks (0 = none)
  rb_u32 dynamic;        // 1: lanes claim further trajectories from *work_next when theirs is finished (ring_depth must be 0)
  rb_u32 resuming;       // 1: this launch continues a call cut short by the watchdog: start from progress[]
  rb_u32* progress;      // [ldn] see RB_PROGRESS_*
  rb_u32 n_launched;     // dynamic: threads of the grid = trajectories assigned statically at the start
  rb_u32* work_next;     // dynamic: [1] trajectories claimed beyond n_launched (zero before the launch)
  rb_u32 bias_hi;        // 0x43300000: high word of the biased-double species form (opaque to the compiler on purpose)
  rb_u32 exp_one;        // 0x3ff: exponent field of 1.0, funnel-shifted into the ziggurat's mantissa word (opaque likewise)
  // Constants the hot loop reads straight from the parameter bank (one LDCU.128 per pair) instead of
  // rebuilding them with two UMOVs each per iteration.
  double bias;           // 2^52 + 2^31
  double one_m_eps;      // 1 - 2^-53
  int byte_sel[4];       // dp4a selectors 1, 1<<8, 1<<16, 1<<24
  rb_u64 save_mask[2];   // specialised kernels: bit s set => species s is sampled
  // event-log mode (nb_steps = 0, src/pyo3_gillespie.rs:209-223): one row per applied reaction
  rb_u32* ev_counts;         // [n_traj] rows of each trajectory (written by the counting pass)
  const rb_u64* ev_offsets;  // [n_traj] first row of each trajectory (read by the writing pass)
  double* ev_times;          // [ev_total] time of every row
  rb_u64 ev_total;           // rows of the whole ensemble = stride of the sample rows in `out`
  rb_u32 ev_single;          // 1: exactly one _advance_one_reaction per trajectory whatever its time, nothing logged
                             // (Gillespie::advance_one_reaction, src/gillespie.rs:270-274)
  const rb_u32* gtab;    // table-driven and large specialised kernels: per-reaction records (+ saved-species list)
  const void* tables;    // table-driven kernel: this batch's RbTables image in global memory
  const rb_u64* pdm;     // partial-propensity kernels: this batch's tables for the reaction choice (pdm.hpp); their
                         // derived constants c_i, K_ij travel in k[]
  int n_species, n_reactions, arith;  // table-driven kernel: copies of the RbTables scalars
  double k[RB_MAX_K];    // specialised kernels: rate constants (kernel parameters may be up to 32 KB on sm_70+)
};

// Dynamic shared memory (bytes) a launch needs: network tables + sample rings.
#define RB_SSA_SMEM_BYTES(net_words, block, ring_depth, n_save) \
  (4u * ((net_words) + ((block) / 32u) * (ring_depth) * (n_save) * 32u))
