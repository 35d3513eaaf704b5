/*
 * ssa_oracle.h -- CPU restatement of rebop's Gillespie direct method.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rebop_b200/ may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs do.  It exists to answer one question: "what would
 * the reference have produced for this network, seed and time grid?".
 *
 * Parity pin: reproduces the reference's only end-to-end golden vector
 * (tests/test_rebop.py:30-36, rng=42 => S=0, I=227, R=773) together with the
 * rate_lma table (src/gillespie.rs:448-472) and test_eval
 * (src/expr.rs:499-516); see tests/test_oracle.py.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  The RNG stack (rand 0.10.2 / rand_distr 0.6.0,
 * Cargo.lock:579-603) is not vendored in the reference tree; it is restated
 * from the published algorithms (SplitMix64 seeding, xoshiro256++, 53-bit
 * uniform, 256-layer ziggurat Exp1).
 */
#ifndef SSA_ORACLE_H
#define SSA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Arithmetic flavour of the propensities and of the reaction choice. */
enum { ORA_ARITH_API = 0,   /* src/gillespie.rs (function API)           */
       ORA_ARITH_MACRO = 1  /* src/gillespie_macro.rs (define_system!)   */ };

/* Expression byte-code (post-order walk of src/expr.rs:9-21 `Expr`). */
enum { ORA_OP_CONST = 0, ORA_OP_SPECIES = 1, ORA_OP_NEG = 2, ORA_OP_ADD = 3,
       ORA_OP_SUB = 4, ORA_OP_MUL = 5, ORA_OP_DIV = 6, ORA_OP_POW = 7,
       ORA_OP_MAX = 8, ORA_OP_MIN = 9, ORA_OP_EXP = 10 };

typedef struct {
  int32_t op;     /* ORA_OP_* */
  int32_t index;  /* species index for ORA_OP_SPECIES */
  double value;   /* constant for ORA_OP_CONST */
} ora_expr_op;

typedef struct ora_network ora_network;

typedef struct {
  uint64_t s[4];
} ora_rng;

/* ---- RNG stack ---------------------------------------------------------- */
void ora_rng_seed(ora_rng* rng, uint64_t seed);   /* SmallRng::seed_from_u64 */
uint64_t ora_rng_next_u64(ora_rng* rng);          /* xoshiro256++ */
double ora_rng_uniform(ora_rng* rng);             /* rng.random::<f64>() */
double ora_rng_exp1(ora_rng* rng);                /* rng.sample(Exp1) */

/* ---- network ------------------------------------------------------------ */
/* `dense` selects which of the reference's two representations is walked
 * (Rate::LMA + Jump::Flat vs Rate::LMASparse + Jump::Sparse); both must give
 * identical results (tests/test_rebop.py:55-65). */
ora_network* ora_network_new(int n_species, int arith, int dense);
void ora_network_free(ora_network* net);
/* Reactant terms (species index, exponent) in evaluation order, jump as a
 * dense difference vector of length n_species. */
int ora_network_add_lma(ora_network* net, double k, const int32_t* term_idx,
                        const int32_t* term_exp, int n_terms, const int64_t* diff);
int ora_network_add_expr(ora_network* net, const ora_expr_op* prog, int n_ops,
                         const int64_t* diff);
int ora_network_nb_species(const ora_network* net);
int ora_network_nb_reactions(const ora_network* net);

/* Propensity of reaction r in state x (Rate::rate, src/gillespie.rs:71-90). */
double ora_rate(const ora_network* net, int r, const int64_t* x);
/* Expr::eval (src/expr.rs:24-38) on a post-order program. */
double ora_expr_eval(const ora_expr_op* prog, int n_ops, const int64_t* x);

/* ---- simulation --------------------------------------------------------- */
typedef struct {
  int64_t* x;       /* [n_species] */
  double t;
  ora_rng rng;
  uint64_t events;  /* applied reactions so far */
} ora_state;

/* Gillespie::advance_until (src/gillespie.rs:315-344) or the macro's
 * advance_until (src/gillespie_macro.rs:98-126), by net->arith. */
void ora_advance_until(const ora_network* net, ora_state* st, double tmax);
/* Gillespie::_advance_one_reaction (src/gillespie.rs:275-297). */
void ora_advance_one_reaction(const ora_network* net, ora_state* st);

/* pyo3 grid loop (src/pyo3_gillespie.rs:197-208): out[(i*n_save + j)] is
 * species save_idx[j] after advance_until(t_i), t_i = tmax*i/nb_steps.
 * times has nb_steps+1 entries.  Returns the number of applied reactions. */
uint64_t ora_run_grid(const ora_network* net, const int64_t* x0, uint64_t seed,
                      double tmax, int nb_steps, const int32_t* save_idx,
                      int n_save, int64_t* out, double* times);

/* nb_steps == 0 path (src/pyo3_gillespie.rs:209-223): one row per event.
 * Writes at most `cap` rows; returns the number of rows the reference would
 * have produced. */
size_t ora_run_events(const ora_network* net, const int64_t* x0, uint64_t seed,
                      double tmax, const int32_t* save_idx, int n_save,
                      int64_t* out, double* times, size_t cap);

/* Ensemble: trajectory n uses seeds[n]; x0 is [n_species] (x0_stride 0) or
 * [n_traj][n_species] (x0_stride n_species).  out is laid out
 * [step][n_save][n_traj] (int32), like the GPU path.  events_per_traj may be
 * NULL.  Trajectories are split statically over `threads` host threads.
 * Returns the total number of applied reactions. */
uint64_t ora_run_batch(const ora_network* net, const int64_t* x0, size_t x0_stride,
                       const uint64_t* seeds, size_t n_traj, double tmax,
                       int nb_steps, const int32_t* save_idx, int n_save,
                       int32_t* out, uint64_t* events_per_traj, int threads);

/* Hand-specialised straight-line restatements of what define_system! expands
 * to for the benchmark systems (the *fast* CPU form of the reference; used as
 * the timed CPU baseline).  name: "vilar", "dimers", "sir".  params/x0 as in
 * the DSL order.  Layout and return value as ora_run_batch.  Returns
 * UINT64_MAX for an unknown name. */
uint64_t ora_run_batch_macro(const char* name, const double* params,
                             const int64_t* x0, const uint64_t* seeds,
                             size_t n_traj, double tmax, int nb_steps,
                             int32_t* out, uint64_t* events_per_traj, int threads);

#ifdef __cplusplus
}
#endif
#endif
