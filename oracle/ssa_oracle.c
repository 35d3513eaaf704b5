/*
 * ssa_oracle.c -- CPU restatement of rebop's Gillespie direct method.
 * TEST INFRASTRUCTURE ONLY (see ssa_oracle.h).  Build with
 *   gcc -O2 -ffp-contract=off   (Rust never contracts a*b+c into an FMA).
 */
#include "ssa_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "zig_tables.h"

/* ======================================================================== */
/* RNG stack: rand 0.10.2 SmallRng (xoshiro256++ on 64-bit targets),         */
/* rand_distr 0.6.0 Exp1.  Call sites: src/gillespie.rs:184,190,285-286,     */
/* 327,332; src/gillespie_macro.rs:83,114,120.                               */
/* ======================================================================== */

/* SmallRng::seed_from_u64: four SplitMix64 outputs fill the state. */
void ora_rng_seed(ora_rng* rng, uint64_t seed) {
  uint64_t st = seed;
  for (int i = 0; i < 4; ++i) {
    st += 0x9e3779b97f4a7c15ULL;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    rng->s[i] = z ^ (z >> 31);
  }
}

static inline uint64_t rotl64(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }

/* xoshiro256++ */
uint64_t ora_rng_next_u64(ora_rng* rng) {
  uint64_t* s = rng->s;
  uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
  uint64_t t = s[1] << 17;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl64(s[3], 45);
  return result;
}

/* StandardUniform for f64: 53 random bits scaled by 2^-53 -> [0,1). */
double ora_rng_uniform(ora_rng* rng) {
  return (double)(ora_rng_next_u64(rng) >> 11) * 0x1.0p-53;
}

/* Exp1: ziggurat with 256 layers (non-symmetric variant). */
double ora_rng_exp1(ora_rng* rng) {
  for (;;) {
    uint64_t bits = ora_rng_next_u64(rng);
    unsigned i = (unsigned)(bits & 0xff);
    /* 52 mantissa bits with exponent 0 -> [1,2), minus (1 - 2^-53) -> (0,1) */
    uint64_t ub = (bits >> 12) | 0x3ff0000000000000ULL;
    double u12;
    memcpy(&u12, &ub, sizeof u12);
    double u = u12 - (1.0 - 0x1.0p-53);
    double x = u * ora_zig_exp_x[i];
    if (x < ora_zig_exp_x[i + 1]) return x;
    if (i == 0) return ORA_ZIG_EXP_R - log(ora_rng_uniform(rng));
    if (ora_zig_exp_f[i + 1] +
            (ora_zig_exp_f[i] - ora_zig_exp_f[i + 1]) * ora_rng_uniform(rng) <
        exp(-x))
      return x;
  }
}

/* ======================================================================== */
/* Network                                                                   */
/* ======================================================================== */

typedef struct {
  int is_expr;
  double k;
  int n_terms;
  int32_t* term_idx; /* evaluation order */
  int32_t* term_exp;
  uint32_t* dense_exp; /* [S]  Rate::LMA */
  int64_t* dense_diff; /* [S]  Jump::Flat */
  int n_jump;          /*      Jump::Sparse (src/gillespie.rs:125-140) */
  int32_t* jump_idx;
  int64_t* jump_diff;
  int n_ops;
  ora_expr_op* prog;
} ora_reaction;

struct ora_network {
  int n_species, n_reactions, cap;
  int arith, dense;
  ora_reaction* rx;
};

ora_network* ora_network_new(int n_species, int arith, int dense) {
  ora_network* net = (ora_network*)calloc(1, sizeof *net);
  net->n_species = n_species;
  net->arith = arith;
  net->dense = dense;
  return net;
}

void ora_network_free(ora_network* net) {
  if (!net) return;
  for (int r = 0; r < net->n_reactions; ++r) {
    ora_reaction* q = &net->rx[r];
    free(q->term_idx); free(q->term_exp); free(q->dense_exp);
    free(q->dense_diff); free(q->jump_idx); free(q->jump_diff); free(q->prog);
  }
  free(net->rx);
  free(net);
}

static ora_reaction* push_reaction(ora_network* net, const int64_t* diff) {
  if (net->n_reactions == net->cap) {
    net->cap = net->cap ? 2 * net->cap : 8;
    net->rx = (ora_reaction*)realloc(net->rx, (size_t)net->cap * sizeof *net->rx);
  }
  ora_reaction* q = &net->rx[net->n_reactions++];
  memset(q, 0, sizeof *q);
  int S = net->n_species;
  q->dense_diff = (int64_t*)calloc((size_t)(S ? S : 1), sizeof(int64_t));
  q->jump_idx = (int32_t*)calloc((size_t)(S ? S : 1), sizeof(int32_t));
  q->jump_diff = (int64_t*)calloc((size_t)(S ? S : 1), sizeof(int64_t));
  for (int s = 0; s < S; ++s) {
    q->dense_diff[s] = diff[s];
    if (diff[s] != 0) { /* Jump::sparse: keep non-zero entries, ascending */
      q->jump_idx[q->n_jump] = s;
      q->jump_diff[q->n_jump++] = diff[s];
    }
  }
  return q;
}

int ora_network_add_lma(ora_network* net, double k, const int32_t* term_idx,
                        const int32_t* term_exp, int n_terms, const int64_t* diff) {
  for (int j = 0; j < n_terms; ++j)
    if (term_idx[j] < 0 || term_idx[j] >= net->n_species) return -1; /* :227-237 */
  ora_reaction* q = push_reaction(net, diff);
  q->k = k;
  q->n_terms = n_terms;
  q->term_idx = (int32_t*)calloc((size_t)(n_terms ? n_terms : 1), sizeof(int32_t));
  q->term_exp = (int32_t*)calloc((size_t)(n_terms ? n_terms : 1), sizeof(int32_t));
  q->dense_exp = (uint32_t*)calloc((size_t)(net->n_species ? net->n_species : 1), sizeof(uint32_t));
  for (int j = 0; j < n_terms; ++j) {
    q->term_idx[j] = term_idx[j];
    q->term_exp[j] = term_exp[j];
    q->dense_exp[term_idx[j]] += (uint32_t)term_exp[j]; /* Rate::dense :44-46 */
  }
  return 0;
}

int ora_network_add_expr(ora_network* net, const ora_expr_op* prog, int n_ops,
                         const int64_t* diff) {
  ora_reaction* q = push_reaction(net, diff);
  q->is_expr = 1;
  q->k = NAN;
  q->n_ops = n_ops;
  q->prog = (ora_expr_op*)calloc((size_t)(n_ops ? n_ops : 1), sizeof *q->prog);
  memcpy(q->prog, prog, (size_t)n_ops * sizeof *prog);
  return 0;
}

int ora_network_nb_species(const ora_network* net) { return net->n_species; }
int ora_network_nb_reactions(const ora_network* net) { return net->n_reactions; }

/* ======================================================================== */
/* Propensities                                                              */
/* ======================================================================== */

/* Expr::eval, src/expr.rs:24-38.  The program is the post-order walk of the
 * tree, so each node sees its operands evaluated exactly as the recursion
 * does (evaluation order cannot change the value: there are no side effects). */
double ora_expr_eval(const ora_expr_op* prog, int n_ops, const int64_t* x) {
  double stack[64];
  int sp = 0;
  for (int i = 0; i < n_ops; ++i) {
    const ora_expr_op* o = &prog[i];
    switch (o->op) {
      case ORA_OP_CONST: stack[sp++] = o->value; break;
      case ORA_OP_SPECIES: stack[sp++] = (double)x[o->index]; break;
      case ORA_OP_NEG: stack[sp - 1] = -stack[sp - 1]; break;
      case ORA_OP_EXP: stack[sp - 1] = exp(stack[sp - 1]); break;
      default: {
        double b = stack[--sp], a = stack[sp - 1], v;
        switch (o->op) {
          case ORA_OP_ADD: v = a + b; break;
          case ORA_OP_SUB: v = a - b; break;
          case ORA_OP_MUL: v = a * b; break;
          case ORA_OP_DIV: v = a / b; break;
          case ORA_OP_POW: v = pow(a, b); break;   /* f64::powf */
          case ORA_OP_MAX: v = fmax(a, b); break;  /* f64::max ignores NaN */
          case ORA_OP_MIN: v = fmin(a, b); break;
          default: v = NAN;
        }
        stack[sp - 1] = v;
      }
    }
  }
  return sp == 1 ? stack[0] : NAN;
}

/* Rate::rate, src/gillespie.rs:71-90. */
static double rate_api(const ora_network* net, const ora_reaction* q, const int64_t* x) {
  if (q->is_expr) return ora_expr_eval(q->prog, q->n_ops, x); /* :88 */
  double acc = q->k;
  if (net->dense) {
    /* Rate::LMA :73-78 -- every species, ascending; factors (n+1-e)..=n */
    for (int s = 0; s < net->n_species; ++s) {
      int64_t n = x[s];
      for (int64_t f = n + 1 - (int64_t)q->dense_exp[s]; f <= n; ++f) acc = acc * (double)f;
    }
  } else {
    /* Rate::LMASparse :79-87 -- (index, exponent) pairs in stored order */
    for (int j = 0; j < q->n_terms; ++j) {
      int64_t n = x[q->term_idx[j]];
      for (int64_t f = n + 1 - (int64_t)q->term_exp[j]; f <= n; ++f) acc = acc * (double)f;
    }
  }
  return acc;
}

/* `$rate * _rate_lma!(n * self.r) * ...`, src/gillespie_macro.rs:106,133-146:
 * each reactant term is an *integer* falling factorial n(n-1)...(n-e+1)
 * (wrapping, as release-mode isize) converted to f64 once. */
static double rate_macro(const ora_reaction* q, const int64_t* x) {
  if (q->is_expr) return ora_expr_eval(q->prog, q->n_ops, x);
  double acc = q->k;
  for (int j = 0; j < q->n_terms; ++j) {
    int64_t n = x[q->term_idx[j]];
    uint64_t prod = (uint64_t)n;
    for (int64_t i = 1; i < (int64_t)q->term_exp[j]; ++i) prod *= (uint64_t)(n - i);
    acc = acc * (double)(int64_t)prod;
  }
  return acc;
}

double ora_rate(const ora_network* net, int r, const int64_t* x) {
  const ora_reaction* q = &net->rx[r];
  return net->arith == ORA_ARITH_MACRO ? rate_macro(q, x) : rate_api(net, q, x);
}

/* Jump::affect, src/gillespie.rs:142-152. */
static void affect(const ora_network* net, const ora_reaction* q, int64_t* x) {
  if (net->dense) {
    for (int s = 0; s < net->n_species; ++s) x[s] += q->dense_diff[s];
  } else {
    for (int j = 0; j < q->n_jump; ++j) x[q->jump_idx[j]] += q->jump_diff[j];
  }
}

/* make_cumrates, src/gillespie.rs:357-364: running sum from literal 0.0. */
static double make_cumrates(const ora_network* net, const int64_t* x, double* cum) {
  double total = 0.0;
  for (int r = 0; r < net->n_reactions; ++r) {
    cum[r] = total + ora_rate(net, r, x);
    total = cum[r];
  }
  return total;
}

/* choose_cumrate_sum, src/gillespie.rs:402-407: count, not search. */
static int choose_cumrate_sum(double chosen, const double* cum, int R) {
  int n = 0;
  for (int r = 0; r < R; ++r) n += (cum[r] < chosen) ? 1 : 0;
  return n;
}

/* ======================================================================== */
/* Simulation                                                                */
/* ======================================================================== */

static void advance_until_api(const ora_network* net, ora_state* st, double tmax, double* cum) {
  const int R = net->n_reactions;
  for (;;) {
    double total = make_cumrates(net, st->x, cum);      /* :319 */
    if (!(0. < total)) { st->t = tmax; return; }        /* :323-326 */
    st->t += ora_rng_exp1(&st->rng) / total;            /* :327 */
    if (st->t > tmax) { st->t = tmax; return; }         /* :328-331 */
    double chosen = total * ora_rng_uniform(&st->rng);  /* :332 */
    int i = choose_cumrate_sum(chosen, cum, R);         /* :336 */
    if (i >= R) i = R - 1; /* unreachable for finite totals (:339); avoids UB */
    affect(net, &net->rx[i], st->x);                    /* :342 */
    st->events++;
  }
}

/* src/gillespie_macro.rs:98-126 with _choice! (:150-171): first match on
 * `rc < carry + r_j`, carry re-associated exactly like total_rate. */
static void advance_until_macro(const ora_network* net, ora_state* st, double tmax, double* rate) {
  const int R = net->n_reactions;
  for (;;) {
    for (int r = 0; r < R; ++r) rate[r] = ora_rate(net, r, st->x); /* :106 */
    double total = 0.;
    for (int r = 0; r < R; ++r) total = total + rate[r];           /* :107 */
    if (!(total > 0.)) { st->t = tmax; return; }                   /* :110-113 */
    st->t += ora_rng_exp1(&st->rng) / total;                       /* :114 */
    if (st->t > tmax) { st->t = tmax; return; }                    /* :115-118 */
    double rc = total * ora_rng_uniform(&st->rng);                 /* :120 */
    double carry = 0.;
    for (int r = 0; r < R; ++r) {                                  /* :121-124 */
      if (rc < carry + rate[r]) {
        affect(net, &net->rx[r], st->x);
        st->events++;
        break;
      }
      carry = carry + rate[r];
    }
  }
}

void ora_advance_until(const ora_network* net, ora_state* st, double tmax) {
  int R = net->n_reactions;
  double stackbuf[64];
  double* buf = R <= 64 ? stackbuf : (double*)malloc((size_t)R * sizeof(double));
  if (net->arith == ORA_ARITH_MACRO) advance_until_macro(net, st, tmax, buf);
  else advance_until_api(net, st, tmax, buf);
  if (buf != stackbuf) free(buf);
}

/* src/gillespie.rs:275-297: exactly one event, no overshoot check;
 * absorbing state => t = +inf. */
void ora_advance_one_reaction(const ora_network* net, ora_state* st) {
  int R = net->n_reactions;
  double stackbuf[64];
  double* cum = R <= 64 ? stackbuf : (double*)malloc((size_t)R * sizeof(double));
  double total = make_cumrates(net, st->x, cum);
  if (!(0. < total)) {
    st->t = INFINITY;
  } else {
    st->t += ora_rng_exp1(&st->rng) / total;
    double chosen = total * ora_rng_uniform(&st->rng);
    int i = choose_cumrate_sum(chosen, cum, R);
    if (i >= R) i = R - 1;
    affect(net, &net->rx[i], st->x);
    st->events++;
  }
  if (cum != stackbuf) free(cum);
}

uint64_t ora_run_grid(const ora_network* net, const int64_t* x0, uint64_t seed,
                      double tmax, int nb_steps, const int32_t* save_idx,
                      int n_save, int64_t* out, double* times) {
  int S = net->n_species;
  ora_state st;
  st.x = (int64_t*)malloc((size_t)(S ? S : 1) * sizeof(int64_t));
  memcpy(st.x, x0, (size_t)S * sizeof(int64_t));
  st.t = 0.;
  st.events = 0;
  ora_rng_seed(&st.rng, seed);
  for (int i = 0; i <= nb_steps; ++i) {
    double t = tmax * (double)i / (double)nb_steps; /* pyo3_gillespie.rs:201 */
    if (times) times[i] = t;
    ora_advance_until(net, &st, t);
    for (int j = 0; j < n_save; ++j) out[(size_t)i * n_save + j] = st.x[save_idx[j]];
  }
  free(st.x);
  return st.events;
}

size_t ora_run_events(const ora_network* net, const int64_t* x0, uint64_t seed,
                      double tmax, const int32_t* save_idx, int n_save,
                      int64_t* out, double* times, size_t cap) {
  int S = net->n_species;
  ora_state st;
  st.x = (int64_t*)malloc((size_t)(S ? S : 1) * sizeof(int64_t));
  memcpy(st.x, x0, (size_t)S * sizeof(int64_t));
  st.t = 0.;
  st.events = 0;
  ora_rng_seed(&st.rng, seed);
  size_t n = 0;
#define ORA_PUSH()                                                            \
  do {                                                                        \
    if (n < cap) {                                                            \
      times[n] = st.t;                                                        \
      for (int j = 0; j < n_save; ++j) out[n * n_save + j] = st.x[save_idx[j]]; \
    }                                                                         \
    ++n;                                                                      \
  } while (0)
  ORA_PUSH();                        /* pyo3_gillespie.rs:212-215 */
  while (st.t < tmax) {              /* :216 */
    ora_advance_one_reaction(net, &st);
    ORA_PUSH();
  }
#undef ORA_PUSH
  free(st.x);
  return n;
}

/* ---- ensembles ---------------------------------------------------------- */

typedef struct {
  const ora_network* net;
  const int64_t* x0;
  size_t x0_stride;
  const uint64_t* seeds;
  size_t n_traj, lo, hi;
  double tmax;
  int nb_steps;
  const int32_t* save_idx;
  int n_save;
  int32_t* out;
  uint64_t* events_per_traj;
  uint64_t events;
  /* macro-specialised */
  int system;
  const double* params;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  const ora_network* net = j->net;
  int S = net->n_species;
  ora_state st;
  st.x = (int64_t*)malloc((size_t)(S ? S : 1) * sizeof(int64_t));
  uint64_t total = 0;
  for (size_t n = j->lo; n < j->hi; ++n) {
    memcpy(st.x, j->x0 + n * j->x0_stride, (size_t)S * sizeof(int64_t));
    st.t = 0.;
    st.events = 0;
    ora_rng_seed(&st.rng, j->seeds[n]);
    for (int i = 0; i <= j->nb_steps; ++i) {
      double t = j->nb_steps ? j->tmax * (double)i / (double)j->nb_steps : j->tmax;
      ora_advance_until(net, &st, t);
      if (j->out)
        for (int s = 0; s < j->n_save; ++s)
          j->out[((size_t)i * j->n_save + s) * j->n_traj + n] = (int32_t)st.x[j->save_idx[s]];
    }
    if (j->events_per_traj) j->events_per_traj[n] = st.events;
    total += st.events;
  }
  free(st.x);
  j->events = total;
  return NULL;
}

static uint64_t run_jobs(batch_job* proto, void* (*fn)(void*), int threads) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > proto->n_traj) threads = proto->n_traj ? (int)proto->n_traj : 1;
  batch_job* jobs = (batch_job*)calloc((size_t)threads, sizeof *jobs);
  pthread_t* tid = (pthread_t*)calloc((size_t)threads, sizeof *tid);
  for (int k = 0; k < threads; ++k) {
    jobs[k] = *proto;
    jobs[k].lo = proto->n_traj * (size_t)k / (size_t)threads;
    jobs[k].hi = proto->n_traj * (size_t)(k + 1) / (size_t)threads;
  }
  for (int k = 1; k < threads; ++k) pthread_create(&tid[k], NULL, fn, &jobs[k]);
  fn(&jobs[0]);
  uint64_t total = jobs[0].events;
  for (int k = 1; k < threads; ++k) {
    pthread_join(tid[k], NULL);
    total += jobs[k].events;
  }
  free(jobs);
  free(tid);
  return total;
}

uint64_t ora_run_batch(const ora_network* net, const int64_t* x0, size_t x0_stride,
                       const uint64_t* seeds, size_t n_traj, double tmax,
                       int nb_steps, const int32_t* save_idx, int n_save,
                       int32_t* out, uint64_t* events_per_traj, int threads) {
  batch_job proto;
  memset(&proto, 0, sizeof proto);
  proto.net = net; proto.x0 = x0; proto.x0_stride = x0_stride; proto.seeds = seeds;
  proto.n_traj = n_traj; proto.tmax = tmax; proto.nb_steps = nb_steps;
  proto.save_idx = save_idx; proto.n_save = n_save; proto.out = out;
  proto.events_per_traj = events_per_traj;
  return run_jobs(&proto, batch_worker, threads);
}

/* ======================================================================== */
/* define_system! expansions written out by hand (src/gillespie_macro.rs:     */
/* 98-126): species are struct fields, every propensity is one straight-line  */
/* expression, the choice is an if/else chain.  This is the form the Rust     */
/* compiler sees for benchmarks/benches/vilar/vilar.rs:6-25,                  */
/* examples/dimers.rs:4-12 and benchmarks/benches/my_benchmark.rs:6-11.       */
/* ======================================================================== */

/* Vilar { Da, Dr, Dpa, Dpr, Ma, Mr, A, R, C }; params in DSL order:
 * aA apA aR apR bA bR dMA dMR dA dR gA gR gC tA tR */
static uint64_t vilar_advance(int64_t* x, double* t, ora_rng* rng, const double* p, double tmax) {
  const double aA = p[0], apA = p[1], aR = p[2], apR = p[3], bA = p[4], bR = p[5],
               dMA = p[6], dMR = p[7], dA = p[8], dR = p[9], gA = p[10], gR = p[11],
               gC = p[12], tA = p[13], tR = p[14];
  int64_t Da = x[0], Dr = x[1], Dpa = x[2], Dpr = x[3], Ma = x[4], Mr = x[5],
          A = x[6], R = x[7], C = x[8];
  uint64_t ev = 0;
#define VILAR_STORE x[0] = Da; x[1] = Dr; x[2] = Dpa; x[3] = Dpr; x[4] = Ma; x[5] = Mr; x[6] = A; x[7] = R; x[8] = C;
  for (;;) {
    double r0 = gA * (double)Da * (double)A;
    double r1 = gR * (double)Dr * (double)A;
    double r2 = tA * (double)Dpa;
    double r3 = tR * (double)Dpr;
    double r4 = aA * (double)Da;
    double r5 = aR * (double)Dr;
    double r6 = apA * (double)Dpa;
    double r7 = apR * (double)Dpr;
    double r8 = bA * (double)Ma;
    double r9 = bR * (double)Mr;
    double r10 = gC * (double)A * (double)R;
    double r11 = dA * (double)C;
    double r12 = dMA * (double)Ma;
    double r13 = dMR * (double)Mr;
    double r14 = dA * (double)A;
    double r15 = dR * (double)R;
    double c0 = 0. + r0, c1 = c0 + r1, c2 = c1 + r2, c3 = c2 + r3, c4 = c3 + r4,
           c5 = c4 + r5, c6 = c5 + r6, c7 = c6 + r7, c8 = c7 + r8, c9 = c8 + r9,
           c10 = c9 + r10, c11 = c10 + r11, c12 = c11 + r12, c13 = c12 + r13,
           c14 = c13 + r14, c15 = c14 + r15;
    double total = c15;
    if (!(total > 0.)) { *t = tmax; VILAR_STORE return ev; }
    *t += ora_rng_exp1(rng) / total;
    if (*t > tmax) { *t = tmax; VILAR_STORE return ev; }
    double rc = total * ora_rng_uniform(rng);
    if (rc < c0) { Da -= 1; A -= 1; Dpa += 1; }
    else if (rc < c1) { Dr -= 1; A -= 1; Dpr += 1; }
    else if (rc < c2) { Dpa -= 1; Da += 1; A += 1; }
    else if (rc < c3) { Dpr -= 1; Dr += 1; A += 1; }
    else if (rc < c4) { Ma += 1; }
    else if (rc < c5) { Mr += 1; }
    else if (rc < c6) { Ma += 1; }
    else if (rc < c7) { Mr += 1; }
    else if (rc < c8) { A += 1; }
    else if (rc < c9) { R += 1; }
    else if (rc < c10) { A -= 1; R -= 1; C += 1; }
    else if (rc < c11) { C -= 1; R += 1; }
    else if (rc < c12) { Ma -= 1; }
    else if (rc < c13) { Mr -= 1; }
    else if (rc < c14) { A -= 1; }
    else if (rc < c15) { R -= 1; }
    else continue; /* no branch taken: nothing applied (:152) */
    ev++;
  }
#undef VILAR_STORE
}

/* Dimers { gene, mRNA, protein, dimer }; params rtx rtl rdi rdm rdp */
static uint64_t dimers_advance(int64_t* x, double* t, ora_rng* rng, const double* p, double tmax) {
  int64_t G = x[0], M = x[1], P = x[2], D = x[3];
  uint64_t ev = 0;
#define DIMERS_STORE x[0] = G; x[1] = M; x[2] = P; x[3] = D;
  for (;;) {
    double r0 = p[0] * (double)G;
    double r1 = p[1] * (double)M;
    double r2 = p[2] * (double)(int64_t)((uint64_t)P * (uint64_t)(P - 1));
    double r3 = p[3] * (double)M;
    double r4 = p[4] * (double)P;
    double c0 = 0. + r0, c1 = c0 + r1, c2 = c1 + r2, c3 = c2 + r3, c4 = c3 + r4;
    double total = c4;
    if (!(total > 0.)) { *t = tmax; DIMERS_STORE return ev; }
    *t += ora_rng_exp1(rng) / total;
    if (*t > tmax) { *t = tmax; DIMERS_STORE return ev; }
    double rc = total * ora_rng_uniform(rng);
    if (rc < c0) { M += 1; }
    else if (rc < c1) { P += 1; }
    else if (rc < c2) { P -= 2; D += 1; }
    else if (rc < c3) { M -= 1; }
    else if (rc < c4) { P -= 1; }
    else continue;
    ev++;
  }
#undef DIMERS_STORE
}

/* SIR { S, I, R }; params r_inf r_heal */
static uint64_t sir_advance(int64_t* x, double* t, ora_rng* rng, const double* p, double tmax) {
  int64_t S = x[0], I = x[1], R = x[2];
  uint64_t ev = 0;
#define SIR_STORE x[0] = S; x[1] = I; x[2] = R;
  for (;;) {
    double r0 = p[0] * (double)S * (double)I;
    double r1 = p[1] * (double)I;
    double c0 = 0. + r0, c1 = c0 + r1;
    double total = c1;
    if (!(total > 0.)) { *t = tmax; SIR_STORE return ev; }
    *t += ora_rng_exp1(rng) / total;
    if (*t > tmax) { *t = tmax; SIR_STORE return ev; }
    double rc = total * ora_rng_uniform(rng);
    if (rc < c0) { S -= 1; I += 1; }
    else if (rc < c1) { I -= 1; R += 1; }
    else continue;
    ev++;
  }
#undef SIR_STORE
}

typedef uint64_t (*macro_advance_fn)(int64_t*, double*, ora_rng*, const double*, double);
static const struct { const char* name; int S; macro_advance_fn fn; } k_systems[] = {
    {"vilar", 9, vilar_advance}, {"dimers", 4, dimers_advance}, {"sir", 3, sir_advance}};

static void* macro_worker(void* arg) {
  batch_job* j = (batch_job*)arg;
  int S = k_systems[j->system].S;
  macro_advance_fn fn = k_systems[j->system].fn;
  int64_t x[16];
  uint64_t total = 0;
  for (size_t n = j->lo; n < j->hi; ++n) {
    memcpy(x, j->x0, (size_t)S * sizeof(int64_t));
    double t = 0.;
    ora_rng rng;
    ora_rng_seed(&rng, j->seeds[n]);
    uint64_t ev = 0;
    for (int i = 0; i <= j->nb_steps; ++i) {
      double ti = j->nb_steps ? j->tmax * (double)i / (double)j->nb_steps : j->tmax;
      ev += fn(x, &t, &rng, j->params, ti);
      if (j->out)
        for (int s = 0; s < S; ++s)
          j->out[((size_t)i * S + s) * j->n_traj + n] = (int32_t)x[s];
    }
    if (j->events_per_traj) j->events_per_traj[n] = ev;
    total += ev;
  }
  j->events = total;
  return NULL;
}

uint64_t ora_run_batch_macro(const char* name, const double* params,
                             const int64_t* x0, const uint64_t* seeds,
                             size_t n_traj, double tmax, int nb_steps,
                             int32_t* out, uint64_t* events_per_traj, int threads) {
  int sys = -1;
  for (size_t k = 0; k < sizeof k_systems / sizeof k_systems[0]; ++k)
    if (strcmp(name, k_systems[k].name) == 0) sys = (int)k;
  if (sys < 0) return UINT64_MAX;
  batch_job proto;
  memset(&proto, 0, sizeof proto);
  proto.system = sys; proto.params = params; proto.x0 = x0; proto.seeds = seeds;
  proto.n_traj = n_traj; proto.tmax = tmax; proto.nb_steps = nb_steps;
  proto.out = out; proto.events_per_traj = events_per_traj;
  return run_jobs(&proto, macro_worker, threads);
}
