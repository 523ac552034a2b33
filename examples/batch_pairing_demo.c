/* batch_pairing_demo.c -- a C caller of the batched engine through include/pbc_b200.h alone
 * (no PBC headers, no Python, no torch): the "host code in C calling sm_100a CUDA through a thin
 * C ABI" path for callers that already hold their elements as element_to_bytes() output.
 *
 *   batch_pairing_demo <param-file> <P.bin> <Q.bin> <E.bin> [n_gpus]
 *
 * P.bin / Q.bin: n G1 / G2 elements in the reference wire format, back to back.  Writes the n GT
 * elements to E.bin and prints one line with the throughput of the timed second call.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/pbc_b200.h"

static unsigned char *slurp(const char *path, size_t *len, int pinned) {
  FILE *f = fopen(path, "rb");
  unsigned char *buf;
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  *len = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  buf = pinned ? pbc_b200_host_alloc(*len ? *len : 1) : malloc(*len + 1);
  if (!buf) { fprintf(stderr, "allocation failed: %s\n", pbc_b200_last_error()); exit(2); }
  if (fread(buf, 1, *len, f) != *len) { fprintf(stderr, "short read on %s\n", path); exit(2); }
  fclose(f);
  return buf;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv) {
  pbc_b200_pairing_t *pairing;
  size_t plen, l1, l2, n;
  unsigned char *param, *P, *Q, *E;
  double t0, t1;
  FILE *out;
  if (argc < 5) { fprintf(stderr, "usage: %s <param> <P.bin> <Q.bin> <E.bin> [n_gpus]\n", argv[0]); return 2; }
  param = slurp(argv[1], &plen, 0);
  if (pbc_b200_pairing_init_set_buf(&pairing, (const char *)param, plen)) {
    fprintf(stderr, "pairing init failed: %s\n", pbc_b200_last_error());
    return 1;
  }
  P = slurp(argv[2], &l1, 1);
  Q = slurp(argv[3], &l2, 1);
  n = l1 / (size_t)pbc_b200_pairing_length_in_bytes_G1(pairing);
  if (n * pbc_b200_pairing_length_in_bytes_G1(pairing) != l1 || n * pbc_b200_pairing_length_in_bytes_G2(pairing) != l2) {
    fprintf(stderr, "input sizes do not describe the same number of elements\n");
    return 2;
  }
  E = pbc_b200_host_alloc(n * pbc_b200_pairing_length_in_bytes_GT(pairing) + 1);
  if (argc > 5 && pbc_b200_set_devices(pairing, atoi(argv[5]))) {
    fprintf(stderr, "set_devices: %s\n", pbc_b200_last_error());
    return 1;
  }
  if (pbc_b200_pairings_apply(pairing, E, P, Q, n)) {           /* warm-up: contexts, workspaces */
    fprintf(stderr, "pairings_apply: %s\n", pbc_b200_last_error());
    return 1;
  }
  t0 = now_s();
  if (pbc_b200_pairings_apply(pairing, E, P, Q, n)) return 1;
  t1 = now_s();
  out = fopen(argv[4], "wb");
  fwrite(E, 1, n * pbc_b200_pairing_length_in_bytes_GT(pairing), out);
  fclose(out);
  printf("{\"demo\": \"batch_pairing\", \"type\": \"%c\", \"n\": %zu, \"seconds\": %.6f, \"pairings_per_s\": %.1f, \"kernel_launches\": %llu}\n",
         pbc_b200_pairing_type(pairing), n, t1 - t0, n / (t1 - t0), (unsigned long long)pbc_b200_kernel_launches());
  pbc_b200_host_free(P); pbc_b200_host_free(Q); pbc_b200_host_free(E);
  pbc_b200_pairing_clear(pairing);
  free(param);
  return 0;
}
