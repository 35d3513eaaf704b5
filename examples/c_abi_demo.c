/*
 * c_abi_demo.c -- the drop-in boundary from plain C99 (no Python, no C++):
 *   SIR of the reference's tests (tests/test_rebop.py:8-12) through include/rebop_b200.h.
 *
 *   gcc -std=c99 -I include examples/c_abi_demo.c -L rebop_b200 -lrebop_b200 -Wl,-rpath,$PWD/rebop_b200 -o c_abi_demo
 *
 * Host-side entry points (network construction, define_system! parsing, code generation) run anywhere; the
 * ensemble itself needs a CUDA device: without one the program reports the engine's error and stops there
 * (exit code 0 with "no device", so that it can run in CPU-only CI).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rebop_b200.h"

#define CHECK(call)                                                                 \
  do {                                                                              \
    int st__ = (call);                                                              \
    if (st__ != REBOP_OK) {                                                         \
      fprintf(stderr, "%s -> %d: %s\n", #call, st__, rebop_b200_last_error());      \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

int main(void) {
  printf("librebop_b200 %s, %d CUDA device(s)\n", rebop_b200_version(), rebop_b200_device_count());

  /* Gillespie::new + add_reaction(Rate::lma(..), ..) (src/lib.rs:122-127) */
  rebop_network* net = NULL;
  CHECK(rebop_network_create(3, REBOP_ARITH_API, &net));
  const uint32_t e_inf[3] = {1, 1, 0}, e_rec[3] = {0, 1, 0};
  const int64_t d_inf[3] = {-1, 1, 0}, d_rec[3] = {0, -1, 1};
  CHECK(rebop_network_add_reaction_lma(net, 1e-4, e_inf, d_inf));
  CHECK(rebop_network_add_reaction_lma(net, 0.01, e_rec, d_rec));
  uint32_t nr = 0;
  CHECK(rebop_network_nb_reactions(net, &nr));
  size_t src_len = 0;
  CHECK(rebop_network_codegen(net, NULL, 0, &src_len));
  printf("SIR: %u reactions, specialised kernel source of %zu bytes\n", nr, src_len);

  /* the shape assert of add_reaction (src/gillespie.rs:229-233, test issue85_oob) comes back as a status */
  const uint32_t bad_index[1] = {7}, one[1] = {1};
  if (rebop_network_add_reaction_lma_sparse(net, 1.0, bad_index, one, 1, d_rec) != REBOP_ERR_OUT_OF_RANGE) {
    fprintf(stderr, "out-of-range species index was not refused\n");
    return 1;
  }

  /* define_system! text -> network in macro arithmetic */
  rebop_system* sys = NULL;
  CHECK(rebop_system_parse("r1 r2; SIR { S, I, R } r_infection: S + I => I + I @ r1\n r_remission: I => R @ r2", &sys));
  const double params[2] = {1e-4, 0.01};
  rebop_network* net_macro = NULL;
  CHECK(rebop_system_network(sys, params, 2, &net_macro));
  int prebuilt = 0;
  CHECK(rebop_network_has_prebuilt(net_macro, &prebuilt));
  printf("define_system SIR: build-time kernel %s\n", prebuilt ? "linked in" : "absent (NVRTC at run time)");

  if (rebop_b200_device_count() == 0) {
    rebop_batch* b = NULL;
    const int64_t x0[3] = {999, 1, 0};
    int st = rebop_batch_create(net, 0, 1, x0, 0, NULL, 0, &b);
    printf("no device: rebop_batch_create -> %d (%s)\n", st, rebop_b200_last_error());
    rebop_system_destroy(sys);
    rebop_network_destroy(net_macro);
    rebop_network_destroy(net);
    return st == REBOP_ERR_CUDA ? 0 : 1;
  }

  /* the reference's golden vector: numpy default_rng(42).integers(2**64-1) = 14276969152011380359 */
  const int64_t x0[3] = {999, 1, 0};
  const uint64_t seed[1] = {14276969152011380359ull};
  rebop_batch* b = NULL;
  CHECK(rebop_batch_create(net, 0, 1, x0, 0, seed, 0, &b));
  int32_t* out = (int32_t*)malloc(251 * 3 * sizeof(int32_t));
  CHECK(rebop_batch_run_grid(b, 250.0, 250, NULL, 0, out));
  uint64_t events = 0;
  CHECK(rebop_batch_events(b, &events, NULL));
  printf("rng=42: S,I,R at t=250 = %d,%d,%d after %llu events (reference: 0,227,773)\n", out[250 * 3 + 0], out[250 * 3 + 1],
         out[250 * 3 + 2], (unsigned long long)events);
  const int ok = out[250 * 3 + 0] == 0 && out[250 * 3 + 1] == 227 && out[250 * 3 + 2] == 773 && events == 1772;
  free(out);
  rebop_batch_destroy(b);
  rebop_system_destroy(sys);
  rebop_network_destroy(net_macro);
  rebop_network_destroy(net);
  return ok ? 0 : 1;
}
