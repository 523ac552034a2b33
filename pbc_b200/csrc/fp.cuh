// fp.cuh -- F_p arithmetic for sm_100a: fixed-limb Montgomery multiplication, all limbs in registers.
//
// Replaces, for the device path, the reference's arith/montfp.c (mont_mul :334-364,
// fp_add/sub/double/halve/neg :220-330).  Same value semantics (x is held as x*R mod p,
// R = 2^(32 N)), different mechanics:
//   * limbs are 32-bit words (N = 16 for the 512-bit Type-A prime, 6 for the 158/159-bit F/D
//     primes); two words = one of the reference's 64-bit limbs, same little-endian order.
//   * the product a*b and the reduction m*p are both accumulated with
//     mad.lo.cc.u32 / madc.hi.cc.u32 carry chains laid out so that every 32x32 product is ONE
//     IMAD.WIDE.U32 with predicate carry-in/out in SASS: products whose low word lands on an even
//     column go to accumulator `ev`, odd columns to `od`; the two are merged once at the end.
//   * coarsely-integrated operand scanning: after each b-limb the running sum is divisible by
//     2^32, the implied one-word shift swaps the roles of ev and od (no data movement).
//   * no zero flag: zero is all-zero limbs (reference: arith/montfp.c:36-39 keeps a flag).
//
// The modulus and its constants live in __constant__ memory so they appear as c[bank][off]
// operands of IMAD and never cost registers.
#pragma once
#include <stdint.h>

namespace pbcb200 {

constexpr int kMaxLimbs = 36;     // type a1 uses 34; 36 keeps every field 16-byte aligned

struct FpConsts {
  uint32_t p[kMaxLimbs];     // modulus
  uint32_t r2[kMaxLimbs];    // R^2 mod p   (to Montgomery form: mont_mul(x, r2))
  uint32_t one[kMaxLimbs];   // R mod p     (Montgomery 1)
  uint32_t pm2[kMaxLimbs];   // p - 2       (Fermat inversion exponent)
  uint32_t np0;              // -p^-1 mod 2^32
  uint32_t nlimbs;
  uint32_t pad[2];
  uint32_t ninv[8];          // -p^-1 mod 2^160 (five-limb fields: the two-product Montgomery reduction fqw_redc_split)
};

__constant__ FpConsts c_fp;   // single translation unit (engine.cu)

// ---------------------------------------------------------------------------------------------
// carry-chain building blocks.  `volatile` keeps the PTX in source order; ptxas tracks CC.CF and
// is free to interleave independent chains when it schedules SASS.
// ---------------------------------------------------------------------------------------------
#define PBC_ASM asm volatile

// acc[0..N) (+)= x[0], x[2], ... , x[N-2] times y, pairs (acc[j],acc[j+1]) receive x[j]*y.
// `x` is indexed with stride 2 starting at x[0]; returns nothing, leaves carry-out in CC.CF.
template <int N>
__device__ __forceinline__ void chain_mad(uint32_t* acc, const uint32_t* x, uint32_t y) {
  PBC_ASM("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
          : "+r"(acc[0]), "+r"(acc[1]) : "r"(x[0]), "r"(y));
#pragma unroll
  for (int j = 2; j < N; j += 2)
    PBC_ASM("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
            : "+r"(acc[j]), "+r"(acc[j + 1]) : "r"(x[j]), "r"(y));
}

// acc[j],acc[j+1] = x[j]*y  (no carries involved)
template <int N>
__device__ __forceinline__ void chain_mul(uint32_t* acc, const uint32_t* x, uint32_t y) {
#pragma unroll
  for (int j = 0; j < N; j += 2)
    PBC_ASM("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;"
            : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(x[j]), "r"(y));
}

// The shifting chain: acc[j],acc[j+1] = x[j]*y + acc[j+2],acc[j+3] (+ incoming CC.CF), the top
// pair adds `top` (the overflow word of the previous round) instead.  Leaves carry-out in CC.CF.
template <int N>
__device__ __forceinline__ void chain_mad_shift(uint32_t* acc, const uint32_t* x, uint32_t y,
                                                uint32_t top) {
#pragma unroll
  for (int j = 0; j < N - 2; j += 2)
    PBC_ASM("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
            : "=r"(acc[j]), "=r"(acc[j + 1])
            : "r"(x[j]), "r"(y), "r"(acc[j + 2]), "r"(acc[j + 3]));
  PBC_ASM("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.cc.u32 %1, %2, %3, %4;"
          : "=r"(acc[N - 2]), "=r"(acc[N - 1]) : "r"(x[N - 2]), "r"(y), "r"(top));
}

__device__ __forceinline__ uint32_t carry_out() {
  uint32_t c;
  PBC_ASM("addc.u32 %0, 0, 0;" : "=r"(c));
  return c;
}

// ---------------------------------------------------------------------------------------------
// One round of the interleaved multiply/reduce.  On entry (first == false) `z` is the
// accumulator whose low word is zero (even-aligned before the implied shift), `w` the
// odd-aligned one, `top` the overflow word above w.  On exit w's low word is zero: the caller swaps.
// FULL = modulus uses all 32N bits (Type A): the running sum needs one more bit than N+1 words.
// ---------------------------------------------------------------------------------------------
template <int N, bool FULL>
__device__ __forceinline__ void mont_round(uint32_t* z, uint32_t* w, uint32_t& top,
                                           const uint32_t* a, uint32_t bi, bool first) {
  if (first) {
    chain_mul<N>(z, a + 1, bi);   // columns 1.. (odd-aligned)
    chain_mul<N>(w, a, bi);       // columns 0.. (even-aligned)
    top = 0;
  } else {
    // shift by one word: w becomes even-aligned as is; z loses its zero word, z[1] is a lone
    // column-0 word, z[2..] become odd-aligned.
    PBC_ASM("add.cc.u32 %0, %0, %1;" : "+r"(w[0]) : "r"(z[1]));
    chain_mad_shift<N>(z, a + 1, bi, top);    // carry of the add enters at column 1
    if (FULL) top = carry_out();
    chain_mad<N>(w, a, bi);
    if (FULL) {
      PBC_ASM("addc.cc.u32 %0, %0, 0;" : "+r"(z[N - 1]));
      PBC_ASM("addc.u32 %0, %0, 0;" : "+r"(top));
    } else {
      PBC_ASM("addc.u32 %0, %0, 0;" : "+r"(z[N - 1]));
    }
  }
  uint32_t m = w[0] * c_fp.np0;
  chain_mad<N>(z, c_fp.p + 1, m);
  if (FULL) {
    uint32_t c = carry_out();
    top += c;
  }
  chain_mad<N>(w, c_fp.p, m);
  if (FULL) {
    PBC_ASM("addc.cc.u32 %0, %0, 0;" : "+r"(z[N - 1]));
    PBC_ASM("addc.u32 %0, %0, 0;" : "+r"(top));
  } else {
    PBC_ASM("addc.u32 %0, %0, 0;" : "+r"(z[N - 1]));
  }
}

// r = (z >> 32) + w (+ top << 32N), then one conditional subtraction of p.  z has zero low word.
template <int N, bool FULL>
__device__ __forceinline__ void mont_finish(uint32_t* r, const uint32_t* z, const uint32_t* w,
                                            uint32_t top) {
  uint32_t t[N];
  PBC_ASM("add.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(z[1]), "r"(w[0]));
#pragma unroll
  for (int k = 1; k < N - 1; k++)
    PBC_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(t[k]) : "r"(z[k + 1]), "r"(w[k]));
  PBC_ASM("addc.cc.u32 %0, %1, 0;" : "=r"(t[N - 1]) : "r"(w[N - 1]));
  if (FULL) PBC_ASM("addc.u32 %0, %0, 0;" : "+r"(top));
  // d = t - p
  uint32_t d[N], borrow;
  PBC_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(d[0]) : "r"(t[0]), "r"(c_fp.p[0]));
#pragma unroll
  for (int k = 1; k < N; k++)
    PBC_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(d[k]) : "r"(t[k]), "r"(c_fp.p[k]));
  PBC_ASM("subc.u32 %0, 0, 0;" : "=r"(borrow));   // 0 if t >= p, 0xffffffff otherwise
  bool use_d = FULL ? (top != 0 || borrow == 0) : (borrow == 0);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = use_d ? d[k] : t[k];
}

// r = a*b/R mod p, inputs and output fully reduced.  r may alias a or b.
template <int N, bool FULL>
__device__ __forceinline__ void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t x[N], y[N], top;
  mont_round<N, FULL>(x, y, top, a, b[0], true);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_round<N, FULL>(y, x, top, a, b[i], false);
    if (i + 1 < N) mont_round<N, FULL>(x, y, top, a, b[i + 1], false);
  }
  // N even: the last round ran as (y, x): x has the zero low word
  mont_finish<N, FULL>(r, x, y, top);
}

// ---------------------------------------------------------------------------------------------
// Product-scanning (column-wise) variant.  Every 32x32 product is one IMAD.WIDE.U32 with
// carry-OUT only (no carry-in: the .X form issues at half rate on sm_100a, measured), the carry
// is absorbed by an IADD3.X on the ALU pipe, so the FMA and ALU pipes alternate.  Two
// independent 96-bit accumulators (a*b and m*p) per column keep two dependency chains in flight.
// ---------------------------------------------------------------------------------------------
#define PBC_MAC3(t0, t1, t2, x, y)                                                           \
  PBC_ASM("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;" \
          : "+r"(t0), "+r"(t1), "+r"(t2) : "r"(x), "r"(y))

template <int N, bool FULL>
__device__ __forceinline__ void mont_mul_ps(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t m[N], t[N];
  uint32_t u0 = 0, u1 = 0, u2 = 0, v0 = 0, v1 = 0, v2 = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int j = 0; j < i; j++) {
      PBC_MAC3(u0, u1, u2, a[j], b[i - j]);
      PBC_MAC3(v0, v1, v2, m[j], c_fp.p[i - j]);
    }
    PBC_MAC3(u0, u1, u2, a[i], b[0]);
    // v += u
    PBC_ASM("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, %5;"
            : "+r"(v0), "+r"(v1), "+r"(v2) : "r"(u0), "r"(u1), "r"(u2));
    m[i] = v0 * c_fp.np0;
    PBC_MAC3(v0, v1, v2, m[i], c_fp.p[0]);   // v0 == 0 now
    v0 = v1; v1 = v2; v2 = 0;
    u0 = 0; u1 = 0; u2 = 0;
  }
#pragma unroll
  for (int i = N; i < 2 * N - 1; i++) {
#pragma unroll
    for (int j = i - N + 1; j < N; j++) {
      PBC_MAC3(u0, u1, u2, a[j], b[i - j]);
      PBC_MAC3(v0, v1, v2, m[j], c_fp.p[i - j]);
    }
    PBC_ASM("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, %5;"
            : "+r"(v0), "+r"(v1), "+r"(v2) : "r"(u0), "r"(u1), "r"(u2));
    t[i - N] = v0;
    v0 = v1; v1 = v2; v2 = 0;
    u0 = 0; u1 = 0; u2 = 0;
  }
  t[N - 1] = v0;
  uint32_t top = v1;
  uint32_t d[N], borrow;
  PBC_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(d[0]) : "r"(t[0]), "r"(c_fp.p[0]));
#pragma unroll
  for (int k = 1; k < N; k++)
    PBC_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(d[k]) : "r"(t[k]), "r"(c_fp.p[k]));
  PBC_ASM("subc.u32 %0, 0, 0;" : "=r"(borrow));
  bool use_d = FULL ? (top != 0 || borrow == 0) : (borrow == 0);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = use_d ? d[k] : t[k];
}

// Squaring, product scanning: cross products a_j a_{i-j} (j < i-j) are accumulated once and
// doubled, the diagonal a_{i/2}^2 is added after.  N(N+1)/2 + N^2 products instead of 2 N^2.
template <int N, bool FULL>
__device__ __forceinline__ void mont_sqr_ps(uint32_t* r, const uint32_t* a) {
  uint32_t m[N], t[N];
  uint32_t u0 = 0, u1 = 0, u2 = 0, v0 = 0, v1 = 0, v2 = 0;
#pragma unroll
  for (int i = 0; i < 2 * N - 1; i++) {
    const int lo = i < N ? 0 : i - N + 1;
#pragma unroll
    for (int j = lo; 2 * j < i; j++) PBC_MAC3(u0, u1, u2, a[j], a[i - j]);
#pragma unroll
    for (int j = lo; j <= (i < N ? i - 1 : N - 1); j++) PBC_MAC3(v0, v1, v2, m[j], c_fp.p[i - j]);
    // u = 2u (+ a_{i/2}^2), v += u
    PBC_ASM("add.cc.u32 %0, %0, %0; addc.cc.u32 %1, %1, %1; addc.u32 %2, %2, %2;"
            : "+r"(u0), "+r"(u1), "+r"(u2));
    if ((i & 1) == 0) PBC_MAC3(u0, u1, u2, a[i / 2], a[i / 2]);
    PBC_ASM("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, %5;"
            : "+r"(v0), "+r"(v1), "+r"(v2) : "r"(u0), "r"(u1), "r"(u2));
    if (i < N) {
      m[i] = v0 * c_fp.np0;
      PBC_MAC3(v0, v1, v2, m[i], c_fp.p[0]);
    } else {
      t[i - N] = v0;
    }
    v0 = v1; v1 = v2; v2 = 0;
    u0 = 0; u1 = 0; u2 = 0;
  }
  t[N - 1] = v0;
  uint32_t top = v1;
  uint32_t d[N], borrow;
  PBC_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(d[0]) : "r"(t[0]), "r"(c_fp.p[0]));
#pragma unroll
  for (int k = 1; k < N; k++)
    PBC_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(d[k]) : "r"(t[k]), "r"(c_fp.p[k]));
  PBC_ASM("subc.u32 %0, 0, 0;" : "=r"(borrow));
  bool use_d = FULL ? (top != 0 || borrow == 0) : (borrow == 0);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = use_d ? d[k] : t[k];
}

// The multiplier the kernels use.  Measured on B200 (profiles/r1_ubench.jsonl): IMAD.WIDE.U32
// issues at 32/clk/SM with or without carries, so the two multipliers execute the same number of
// multiplier-pipe slots; operand scanning has ~25% fewer ALU instructions and wins for a*b, product
// scanning wins for a^2 (408 instead of 528 IMAD.WIDE).  PBC_MULT_IMPL: 2 = that hybrid (default),
// 1 = product scanning for both, 0 = operand scanning for both (A/B builds in profiles/).
#ifndef PBC_MULT_IMPL
#define PBC_MULT_IMPL 2
#endif
template <int N, bool FULL>
__device__ __forceinline__ void fp_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if PBC_MULT_IMPL == 1
  mont_mul_ps<N, FULL>(r, a, b);
#else
  mont_mul<N, FULL>(r, a, b);
#endif
}
template <int N, bool FULL>
__device__ __forceinline__ void fp_sqr(uint32_t* r, const uint32_t* a) {
#if PBC_MULT_IMPL >= 1
  mont_sqr_ps<N, FULL>(r, a);
#else
  mont_mul<N, FULL>(r, a, a);
#endif
}

// ---------------------------------------------------------------------------------------------
// additive operations (arith/montfp.c:220-330), branch-free conditional correction
// ---------------------------------------------------------------------------------------------
template <int N, bool FULL>
__device__ __forceinline__ void fp_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t t[N], d[N], c = 0, borrow;
  PBC_ASM("add.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(a[0]), "r"(b[0]));
#pragma unroll
  for (int k = 1; k < N; k++)
    PBC_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(t[k]) : "r"(a[k]), "r"(b[k]));
  if (FULL) c = carry_out();
  PBC_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(d[0]) : "r"(t[0]), "r"(c_fp.p[0]));
#pragma unroll
  for (int k = 1; k < N; k++)
    PBC_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(d[k]) : "r"(t[k]), "r"(c_fp.p[k]));
  PBC_ASM("subc.u32 %0, 0, 0;" : "=r"(borrow));
  bool use_d = FULL ? (c != 0 || borrow == 0) : (borrow == 0);
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = use_d ? d[k] : t[k];
}

template <int N>
__device__ __forceinline__ void fp_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t t[N], borrow;
  PBC_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(a[0]), "r"(b[0]));
#pragma unroll
  for (int k = 1; k < N; k++)
    PBC_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(t[k]) : "r"(a[k]), "r"(b[k]));
  PBC_ASM("subc.u32 %0, 0, 0;" : "=r"(borrow));   // all-ones when a < b
  // add back p & mask
  PBC_ASM("add.cc.u32 %0, %1, %2;" : "=r"(r[0]) : "r"(t[0]), "r"(c_fp.p[0] & borrow));
#pragma unroll
  for (int k = 1; k < N; k++)
    PBC_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(r[k]) : "r"(t[k]), "r"(c_fp.p[k] & borrow));
}

template <int N>
__device__ __forceinline__ bool fp_is_zero(const uint32_t* a) {
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < N; k++) o |= a[k];
  return o == 0;
}

template <int N>
__device__ __forceinline__ void fp_neg(uint32_t* r, const uint32_t* a) {
  // p - a, and 0 stays 0
  uint32_t mask = fp_is_zero<N>(a) ? 0u : 0xffffffffu;
  uint32_t t[N];
  PBC_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(c_fp.p[0]), "r"(a[0]));
#pragma unroll
  for (int k = 1; k < N; k++)
    PBC_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(t[k]) : "r"(c_fp.p[k]), "r"(a[k]));
#pragma unroll
  for (int k = 0; k < N; k++) r[k] = t[k] & mask;
}

// r = a/2 mod p  (arith/montfp.c:282-301)
template <int N, bool FULL>
__device__ __forceinline__ void fp_halve(uint32_t* r, const uint32_t* a) {
  uint32_t mask = (a[0] & 1u) ? 0xffffffffu : 0u;
  uint32_t t[N], c;
  PBC_ASM("add.cc.u32 %0, %1, %2;" : "=r"(t[0]) : "r"(a[0]), "r"(c_fp.p[0] & mask));
#pragma unroll
  for (int k = 1; k < N; k++)
    PBC_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(t[k]) : "r"(a[k]), "r"(c_fp.p[k] & mask));
  c = carry_out();
#pragma unroll
  for (int k = 0; k < N - 1; k++) r[k] = __funnelshift_r(t[k], t[k + 1], 1);
  r[N - 1] = __funnelshift_r(t[N - 1], c, 1);
}

template <int N>
__device__ __forceinline__ bool fp_eq(const uint32_t* a, const uint32_t* b) {
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < N; k++) o |= a[k] ^ b[k];
  return o == 0;
}

}  // namespace pbcb200
