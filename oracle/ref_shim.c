/* ref_shim.c -- byte-buffer C entry points over the UNMODIFIED reference library.
 *
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline).  Compiled by oracle/Makefile together with
 * the reference's own sources (taken where they lie under /root/reference, never copied) into
 * oracle/_ref/libpbcref.so.  Everything here goes through the reference's public API
 * (pbc.h: pairing_init_set_buf, element_from_bytes, element_pairing, element_prod_pairing,
 * pairing_pp_*, element_pow_zn, element_to_bytes) so the bytes it returns are the reference's.
 * All buffers use the reference wire format (element_to_bytes / element_from_bytes).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <pbc.h>

typedef struct {
  pairing_t pairing;
  int g1_len, g2_len, gt_len, zr_len;
} pbcref_t;

static double now_s(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}

void *pbcref_open(const char *param, size_t len) {
  pbcref_t *h = calloc(1, sizeof(*h));
  if (pairing_init_set_buf(h->pairing, param, len)) { free(h); return NULL; }
  element_t e;
  element_init_G1(e, h->pairing); h->g1_len = element_length_in_bytes(e); element_clear(e);
  element_init_G2(e, h->pairing); h->g2_len = element_length_in_bytes(e); element_clear(e);
  element_init_GT(e, h->pairing); h->gt_len = element_length_in_bytes(e); element_clear(e);
  element_init_Zr(e, h->pairing); h->zr_len = element_length_in_bytes(e); element_clear(e);
  return h;
}

void pbcref_close(void *hv) {
  pbcref_t *h = hv;
  pairing_clear(h->pairing);
  free(h);
}

void pbcref_sizes(void *hv, int *out4) {
  pbcref_t *h = hv;
  out4[0] = h->g1_len; out4[1] = h->g2_len; out4[2] = h->gt_len; out4[3] = h->zr_len;
}

void pbcref_seed(unsigned long seed) { pbc_random_set_deterministic((unsigned int)seed); }

static void init_group(element_t e, pbcref_t *h, int group) {
  switch (group) {
    case 1: element_init_G1(e, h->pairing); break;
    case 2: element_init_G2(e, h->pairing); break;
    case 3: element_init_GT(e, h->pairing); break;
    default: element_init_Zr(e, h->pairing); break;
  }
}

static int group_len(pbcref_t *h, int group) {
  return group == 1 ? h->g1_len : group == 2 ? h->g2_len : group == 3 ? h->gt_len : h->zr_len;
}

/* n independent element_random() draws from group (1=G1, 2=G2, 3=GT, 0=Zr). */
void pbcref_random(void *hv, int group, unsigned char *out, size_t n) {
  pbcref_t *h = hv;
  element_t e;
  init_group(e, h, group);
  int len = group_len(h, group);
  for (size_t i = 0; i < n; i++) {
    element_random(e);
    element_to_bytes(out + i * len, e);
  }
  element_clear(e);
}

/* Cheap large batches: X_0 random, X_{i+1} = X_i + G, G random. */
void pbcref_walk(void *hv, int group, unsigned char *out, size_t n) {
  pbcref_t *h = hv;
  element_t e, g;
  init_group(e, h, group);
  init_group(g, h, group);
  int len = group_len(h, group);
  element_random(e);
  element_random(g);
  for (size_t i = 0; i < n; i++) {
    element_to_bytes(out + i * len, e);
    element_add(e, e, g);
  }
  element_clear(e);
  element_clear(g);
}

/* out[i] = e(P[i], Q[i]); returns elapsed seconds of the pairing calls alone. */
double pbcref_pairing(void *hv, const unsigned char *P, const unsigned char *Q,
                      unsigned char *out, size_t n) {
  pbcref_t *h = hv;
  element_t p, q, r;
  element_init_G1(p, h->pairing);
  element_init_G2(q, h->pairing);
  element_init_GT(r, h->pairing);
  double t = 0;
  for (size_t i = 0; i < n; i++) {
    element_from_bytes(p, (unsigned char *)P + i * h->g1_len);
    element_from_bytes(q, (unsigned char *)Q + i * h->g2_len);
    double t0 = now_s();
    element_pairing(r, p, q);
    t += now_s() - t0;
    element_to_bytes(out + i * h->gt_len, r);
  }
  element_clear(p); element_clear(q); element_clear(r);
  return t;
}

/* out[i] = prod_{j<k} e(P[i*k+j], Q[i*k+j]) via element_prod_pairing. */
double pbcref_prod_pairing(void *hv, const unsigned char *P, const unsigned char *Q,
                           unsigned char *out, size_t k, size_t n_out) {
  pbcref_t *h = hv;
  element_t *p = malloc(k * sizeof(element_t)), *q = malloc(k * sizeof(element_t)), r;
  for (size_t j = 0; j < k; j++) { element_init_G1(p[j], h->pairing); element_init_G2(q[j], h->pairing); }
  element_init_GT(r, h->pairing);
  double t = 0;
  for (size_t i = 0; i < n_out; i++) {
    for (size_t j = 0; j < k; j++) {
      element_from_bytes(p[j], (unsigned char *)P + (i * k + j) * h->g1_len);
      element_from_bytes(q[j], (unsigned char *)Q + (i * k + j) * h->g2_len);
    }
    double t0 = now_s();
    element_prod_pairing(r, p, q, (int)k);
    t += now_s() - t0;
    element_to_bytes(out + i * h->gt_len, r);
  }
  for (size_t j = 0; j < k; j++) { element_clear(p[j]); element_clear(q[j]); }
  element_clear(r); free(p); free(q);
  return t;
}

/* Fixed first argument: out[i] = e(P, Q[i]) through pairing_pp_init / pairing_pp_apply. */
double pbcref_pp_pairing(void *hv, const unsigned char *P, const unsigned char *Q,
                         unsigned char *out, size_t n) {
  pbcref_t *h = hv;
  element_t p, q, r;
  pairing_pp_t pp;
  element_init_G1(p, h->pairing);
  element_init_G2(q, h->pairing);
  element_init_GT(r, h->pairing);
  element_from_bytes(p, (unsigned char *)P);
  pairing_pp_init(pp, p, h->pairing);
  double t = 0;
  for (size_t i = 0; i < n; i++) {
    element_from_bytes(q, (unsigned char *)Q + i * h->g2_len);
    double t0 = now_s();
    pairing_pp_apply(r, q, pp);
    t += now_s() - t0;
    element_to_bytes(out + i * h->gt_len, r);
  }
  pairing_pp_clear(pp);
  element_clear(p); element_clear(q); element_clear(r);
  return t;
}

/* out[i] = in[i]^k[i]  (element_pow_zn; additive groups: scalar multiple). */
void pbcref_pow_zn(void *hv, int group, const unsigned char *in, const unsigned char *k,
                   unsigned char *out, size_t n) {
  pbcref_t *h = hv;
  element_t e, z;
  init_group(e, h, group);
  element_init_Zr(z, h->pairing);
  int len = group_len(h, group);
  for (size_t i = 0; i < n; i++) {
    element_from_bytes(e, (unsigned char *)in + i * len);
    element_from_bytes(z, (unsigned char *)k + i * h->zr_len);
    element_pow_zn(e, e, z);
    element_to_bytes(out + i * len, e);
  }
  element_clear(e); element_clear(z);
}

/* out[i] = element_from_hash(data + i*len, len) in the group (include/pbc_field.h:202-212;
 * G1/G2: ecc/curve.c:455-482 curve_from_hash). */
void pbcref_from_hash(void *hv, int group, const unsigned char *data, int len, unsigned char *out, size_t n) {
  pbcref_t *h = hv;
  element_t e;
  init_group(e, h, group);
  int elen = group_len(h, group);
  for (size_t i = 0; i < n; i++) {
    element_from_hash(e, (void *)(data + i * (size_t)len), len);
    element_to_bytes(out + i * elen, e);
  }
  element_clear(e);
}

/* element_to_bytes_compressed / element_from_bytes_compressed (ecc/curve.c:762-813) on G1 or G2:
 * compress: wire -> x || sign byte;  decompress: x || sign byte -> wire. */
int pbcref_compressed_len(void *hv, int group) {
  pbcref_t *h = hv;
  element_t e;
  init_group(e, h, group);
  int n = element_length_in_bytes_compressed(e);
  element_clear(e);
  return n;
}
void pbcref_compress(void *hv, int group, const unsigned char *in, unsigned char *out, size_t n) {
  pbcref_t *h = hv;
  element_t e;
  init_group(e, h, group);
  int len = group_len(h, group), clen = element_length_in_bytes_compressed(e);
  for (size_t i = 0; i < n; i++) {
    element_from_bytes(e, (unsigned char *)in + i * len);
    element_to_bytes_compressed(out + i * clen, e);
  }
  element_clear(e);
}
void pbcref_decompress(void *hv, int group, const unsigned char *in, unsigned char *out, size_t n) {
  pbcref_t *h = hv;
  element_t e;
  init_group(e, h, group);
  int len = group_len(h, group), clen = element_length_in_bytes_compressed(e);
  for (size_t i = 0; i < n; i++) {
    element_from_bytes_compressed(e, (unsigned char *)in + i * clen);
    element_to_bytes(out + i * len, e);
  }
  element_clear(e);
}

/* out[i] = a[i] * b[i] in the group (additive groups: a+b). */
void pbcref_mul(void *hv, int group, const unsigned char *a, const unsigned char *b,
                unsigned char *out, size_t n) {
  pbcref_t *h = hv;
  element_t x, y;
  init_group(x, h, group);
  init_group(y, h, group);
  int len = group_len(h, group);
  for (size_t i = 0; i < n; i++) {
    element_from_bytes(x, (unsigned char *)a + i * len);
    element_from_bytes(y, (unsigned char *)b + i * len);
    element_mul(x, x, y);
    element_to_bytes(out + i * len, x);
  }
  element_clear(x); element_clear(y);
}

/* Parse a decimal "[x, y]" style string (element_set_str) and return wire bytes. */
int pbcref_from_str(void *hv, int group, const char *s, unsigned char *out) {
  pbcref_t *h = hv;
  element_t e;
  init_group(e, h, group);
  int r = element_set_str(e, s, 10);
  element_to_bytes(out, e);
  element_clear(e);
  return r;
}

/* 1 if the bytes decode to the identity (for curve groups: off-curve input decodes to O). */
int pbcref_is_identity(void *hv, int group, const unsigned char *in) {
  pbcref_t *h = hv;
  element_t e;
  init_group(e, h, group);
  element_from_bytes(e, (unsigned char *)in);
  int r = group == 0 ? element_is0(e) : element_is1(e);
  element_clear(e);
  return r;
}
