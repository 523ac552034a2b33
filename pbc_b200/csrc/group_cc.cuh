// group_cc.cuh -- the operations either side of the Type F / Type D pairings:
//   element_pow_zn on G1 = E(F_q)  (ecc/curve.c:455-482 -> arith/field.c:113-126 windowed power over
//       curve_double / curve_mul, one inversion per affine operation)
//   element_pow_zn on GT (ecc/pairing.c:199-231 -> windowed power in F_q^12 / F_q^6)
// Same values, different route: left-to-right double-and-add on Jacobian coordinates (the formulas
// of miller_cc.cuh without the lines) with one Fermat inversion at the end, and square-and-multiply
// on the tower routines of pairing_f.cuh / pairing_d.cuh.  G2 of these types (curves over F_q^2 and
// F_q^3) is not built yet.
#pragma once
#include "group_a.cuh"
#include "pairing_d.cuh"
#include "pairing_f.cuh"

namespace pbcb200 {

// out[i] = k[i] * in[i] on y^2 = x^3 + A x + B over the five-limb field.  O -> zero bytes.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, PBC_CC_MINBLOCKS)
k_cc_g1_mul(const uint8_t* __restrict__ P, const uint8_t* __restrict__ K, uint8_t* __restrict__ out,
            size_t n) {
  size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (idx >= n) return;
  Fq xP, yP, X, Y, Z, Z2, M, Y2, t, u, H, R;
  fq_from_wire(xP, P + idx * (2 * kWS));
  fq_from_wire(yP, P + idx * (2 * kWS) + kWS);
  bool ok = cc_on_curve(xP, yP);
  uint32_t k[5];
  zr_from_wire(k, K + idx * kWZ);
  int top = zr_top_bit(k);
  X = xP;
  Y = yP;
  fq_one(Z);
  for (int j = top - 1; j >= 0; j--) {
    // V = 2V
    fq_sqr(Z2, Z);
    fq_sqr(t, X);
    fq_dbl(M, t);
    fq_add(M, M, t);
    if (!c_cc.a_is_zero) {
      fq_sqr(u, Z2);
      fq_set(t, c_cc.A);
      fq_mul(u, u, t);
      fq_add(M, M, u);
    }
    fq_sqr(Y2, Y);
    fq_mul(u, Y, Z);
    fq_dbl(Z, u);
    fq_mul(t, X, Y2);
    fq_dbl(t, t);
    fq_dbl(t, t);
    fq_sqr(X, M);
    fq_sub(X, X, t);
    fq_sub(X, X, t);
    fq_sqr(Y2, Y2);
    fq_dbl(Y2, Y2);
    fq_dbl(Y2, Y2);
    fq_dbl(Y2, Y2);
    fq_sub(t, t, X);
    fq_mul(Y, M, t);
    fq_sub(Y, Y, Y2);
    if ((k[j >> 5] >> (j & 31)) & 1u) {
      // V = V + P (mixed)
      fq_sqr(Z2, Z);
      fq_mul(t, Z2, Z);
      fq_mul(H, xP, Z2);
      fq_sub(H, H, X);
      fq_mul(R, yP, t);
      fq_sub(R, R, Y);
      fq_mul(Z, H, Z);
      fq_sqr(t, H);
      fq_mul(u, t, H);
      fq_mul(t, t, X);
      fq_sqr(X, R);
      fq_sub(X, X, u);
      fq_sub(X, X, t);
      fq_sub(X, X, t);
      fq_sub(t, t, X);
      fq_mul(t, t, R);
      fq_mul(u, u, Y);
      fq_sub(Y, t, u);
    }
  }
  bool inf = !ok || top < 0 || fq_is_zero(Z);
  fq_inv(&t, &Z);
  fq_sqr(u, t);
  fq_mul(X, X, u);
  fq_mul(u, u, t);
  fq_mul(Y, Y, u);
  if (inf) { fq_zero(X); fq_zero(Y); }
  fq_to_wire(out + idx * (2 * kWS), X);
  fq_to_wire(out + idx * (2 * kWS) + kWS, Y);
}

// out[i] = in[i]^k[i] in GT (type f: 240-byte F_q^12 elements)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, PBC_CC_MINBLOCKS)
k_f_gt_pow(const uint8_t* __restrict__ G, const uint8_t* __restrict__ K, uint8_t* __restrict__ out,
           size_t n) {
  size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (idx >= n) return;
  F12 base, acc;
  uint32_t k[5];
  f12_from_wire(base, G + idx * (12 * kWS));
  f12_to_internal(base);
  zr_from_wire(k, K + idx * kWZ);
  int top = zr_top_bit(k);
  acc = base;
  for (int j = top - 1; j >= 0; j--) {
    f12_sqr(&acc);
    if ((k[j >> 5] >> (j & 31)) & 1u) f12_mul(&acc, &acc, &base);
  }
  if (top < 0) f12_one(acc);
  f12_to_reference(acc);
  f12_to_wire(out + idx * (12 * kWS), acc);
}

// type d: 120-byte F_q^6 elements
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, PBC_CC_MINBLOCKS)
k_d_gt_pow(const uint8_t* __restrict__ G, const uint8_t* __restrict__ K, uint8_t* __restrict__ out,
           size_t n) {
  size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (idx >= n) return;
  F6D base, acc;
  uint32_t k[5];
  f6d_from_wire(base, G + idx * (6 * kWS));
  zr_from_wire(k, K + idx * kWZ);
  int top = zr_top_bit(k);
  acc = base;
  for (int j = top - 1; j >= 0; j--) {
    f6d_sqr(&acc);
    if ((k[j >> 5] >> (j & 31)) & 1u) f6d_mul(&acc, &acc, &base);
  }
  if (top < 0) f6d_one(acc);
  f6d_to_wire(out + idx * (6 * kWS), acc);
}

}  // namespace pbcb200
