// pairing_f_pair.cuh -- Type F Miller loop on the shared-memory slot machine, TWO lanes per pairing.
//
// Same function, same slots, same arithmetic as k_f_miller_s (pairing_f_slots.cuh; device replacement for
// cc_miller_no_denom, ecc/f_param.c:97-248).  What changes is who executes it.  ncu on k_f_miller_s: shared
// memory (720 B of slots per pairing) allows 256 pairings per SM, i.e. two warps per scheduler when a thread owns a
// pairing; a warp needs 4.5 cycles per instruction (fixed-latency dependencies of the carry chains: stall_wait 2.0
// per issue), so the schedulers issue 44 % of the time and the multiplier pipe idles a third of it.  The slots
// already live in shared memory, where any lane can read them -- so the two adjacent lanes (2j, 2j+1) of a warp share
// pairing j: 512 threads = four warps per scheduler per SM for the same shared memory, half the work each:
//
//   * the F_q^12 squaring's two F_q^6 products (A B and (A + B)(A + yB)) are independent: one per lane;
//   * the sparse line product computes coefficients 0..2 on lane 0 and 3..5 on lane 1;
//   * F_q^2 additions, xi-multiples and scalings: real part on lane 0, imaginary part on lane 1;
//   * the F_q products of the point arithmetic go out in independent pairs, one per lane;
//   * single-slot additions are executed by both lanes (same value stored twice).
//
// Both lanes run ONE instruction stream -- the lane parity only selects slot numbers and signs -- so the warp
// never diverges.  A routine reads its operands, works in registers, then __syncwarp(), stores, __syncwarp():
// the first barrier keeps a lane from overwriting what its partner has not read yet, the second publishes the
// results.  Decoding / validation (k_f_prep) stays one thread per pairing in its own kernel, so the loop kernel
// can be held to 128 registers.
#pragma once
#include "pairing_f_slots.cuh"

namespace pbcb200 {

// 44 slots = 880 B per pairing (two 128-pairing blocks per SM): the 36 of k_f_miller_s + 8 of scratch.  During
// the F_q^12 squaring the line's slots are dead, so slots 28..43 are its scratch: A + yB (6), the second
// product (6), xi B_2 and xi (A + yB)_2 (2 + 2).
constexpr int kFPSlots = 44;
// Slot map of the pair kernel (0..26 as k_f_miller_s: value, scratch area, X, Y, Z).  The line's slots are laid
// out so that the 16 scratch slots of the squaring are 28..43 and so that the bank rule below works out:
//   27 c | 28..33: L3, L4, (2 spare) | 34..39: xi L3, xi L4, (2 spare) | 40, 41 | 42, 43
//   squaring scratch:  28..33 = A + yB,  34..39 = the second product,  40, 41 = xi B_2,  42, 43 = xi (A + yB)_2
enum FPSlotMap { fpC = 27, fpL3 = 28, fpL4 = 30, fpXL3 = 34, fpXL4 = 36, fpScratch = 28 };
// PBC_FP_SWIZZLE = 1: word w of slot s of pairing p lives at row (s * 5 + w), column p ^ (16 sigma(s)).  Two lanes of
// a pair that read different slots in the same instruction would otherwise hit the same 16 banks with two addresses each
// (rows are 128 words apart; ncu: 1.2 G two-way conflicts per launch).  sigma(s) = (s & 1) ^ alpha(s), alpha = 0 on the
// low halves of the two 12-slot areas and on 28..33, 40, 41; 1 on the high halves and on 34..39, 42, 43: every pair of
// slots the big routines touch together (real / imaginary coordinate; A and the second area's high half; B and A + yB;
// L and xi L; coefficient k and k + 3 of the line product; the two products' outputs) then differs in sigma.
// Measured (profiles/r2_variants_pairsw.jsonl): 57.5 ms with the swizzle against 55.8 ms without -- the conflicts are
// not what holds this kernel back and the extra address arithmetic costs more than they do.  Off; the simulator test
// builds it on.
#ifndef PBC_FP_SWIZZLE
#define PBC_FP_SWIZZLE 0
#endif
constexpr uint64_t fp_sigma_mask() {
  uint64_t m = 0;
  for (int s = 0; s < kFPSlots; s++) {
    int alpha = 0;
    if (s < 24) alpha = (s % 12) >= 6;
    else if (s >= 34 && s <= 39) alpha = 1;
    else if (s >= 42) alpha = 1;
    if (((s & 1) ^ alpha) != 0) m |= (uint64_t)1 << s;
  }
  return m;
}
constexpr uint64_t kFPSigma = fp_sigma_mask();

template <int BP>                     // pairings per block; the block has 2 BP threads
struct FP {
  static __device__ __forceinline__ int half() { return (int)(threadIdx.x & 1u); }
  static __device__ __forceinline__ uint32_t* base(int s) {
    uint32_t col = threadIdx.x >> 1;
    if (PBC_FP_SWIZZLE) col ^= (uint32_t)((kFPSigma >> s) & 1u) << 4;
    return reinterpret_cast<uint32_t*>(pbc_smem) + col;
  }
  static __device__ __forceinline__ void ld(Fq& r, int s) {
    const uint32_t* b = base(s) + s * (kNS * BP);
#pragma unroll
    for (int k = 0; k < kNS; k++) r.v[k] = b[k * BP];
  }
  static __device__ __forceinline__ void st(int s, const Fq& r) {
    uint32_t* b = base(s) + s * (kNS * BP);
#pragma unroll
    for (int k = 0; k < kNS; k++) b[k * BP] = r.v[k];
  }
  static __device__ __forceinline__ void sync() { __syncwarp(); }
  // store after the partner has read, publish
  static __device__ __forceinline__ void put(int s, const Fq& r) { sync(); st(s, r); sync(); }

  // ---- F_q: one product per lane.  lane 0: d0 = a0 b0, lane 1: d1 = a1 b1 (a lone product: the same triple twice)
  static __device__ __noinline__ void qmulh(int d0, int a0, int b0, int d1, int a1, int b1) {
    const bool hh = half() != 0;
    const int d = hh ? d1 : d0, a = hh ? a1 : a0, b = hh ? b1 : b0;
    Fq x, y;
    ld(x, a); ld(y, b);
    fq_mul_os(x, x, y);
    put(d, x);
  }
  // single-slot linear operations: both lanes compute and store the same value
  static __device__ __noinline__ void qadd(int d, int a, int b) { Fq x, y; ld(x, a); ld(y, b); fq_add(x, x, y); put(d, x); }
  static __device__ __noinline__ void qsub(int d, int a, int b) { Fq x, y; ld(x, a); ld(y, b); fq_sub(x, x, y); put(d, x); }
  static __device__ __noinline__ void qdbl(int d, int a, int k = 1) {
    Fq x;
    ld(x, a);
    for (int i = 0; i < k; i++) fq_dbl(x, x);
    put(d, x);
  }
  static __device__ __noinline__ void qneg(int d, int a) { Fq x; ld(x, a); fq_neg(x, x); put(d, x); }
  static __device__ __noinline__ void qcopy(int d, int a) { Fq x; ld(x, a); put(d, x); }
  static __device__ __noinline__ void qldg(int d, const uint32_t* g, int e, size_t n, bool neg) {
    Fq x;
#pragma unroll
    for (int k = 0; k < kNS; k++) x.v[k] = g[((size_t)e * kNS + k) * n];
    if (neg) fq_neg(x, x);
    put(d, x);
  }
  static __device__ __noinline__ void qldc(int d, const uint32_t* c) {
    Fq x;
#pragma unroll
    for (int k = 0; k < kNS; k++) x.v[k] = c[k];
    put(d, x);
  }
  static __device__ __noinline__ void qzero(int d) { Fq x; fq_zero(x); put(d, x); }
  // (d, d+1) <- (global F_q^2 element at e, e+1) * slot a: lane h makes coordinate h
  static __device__ __noinline__ void f2scale_g(int d, const uint32_t* g, int e, size_t n, int a) {
    const int hh = half();
    Fq x, y;
#pragma unroll
    for (int k = 0; k < kNS; k++) y.v[k] = g[((size_t)(e + hh) * kNS + k) * n];
    ld(x, a);
    fq_mul_os(y, y, x);
    put(d + hh, y);
  }

  // ---- F_q^2: lane h owns coordinate h ----
  // coordinate h of xi' (y0 + y1 i), xi' = a + b i:  a y0 - b y1 (h = 0),  a y1 + b y0 (h = 1);  u = y_h, w = y_(1-h)
  static __device__ __forceinline__ void xi_part(Fq& r, const Fq& u, const Fq& w, bool hh) {
    const uint32_t a = c_f.xi_a, b = c_f.xi_b;
    Fq u2, u4, w2, w4, p, t, nt;
    if (a & 6u) fq_dbl(u2, u);
    if (a & 4u) fq_dbl(u4, u2);
    if (b & 6u) fq_dbl(w2, w);
    if (b & 4u) fq_dbl(w4, w2);
    fq_small_combo(p, a, u, u2, u4);         // a u
    fq_small_combo(t, b, w, w2, w4);         // b w
    fq_neg(nt, t);
#pragma unroll
    for (int k = 0; k < kNS; k++) t.v[k] = hh ? t.v[k] : nt.v[k];
    fq_add(r, p, t);
  }
  // d = a + b (MODE 0), a - b (1), 2 a (2), a + xi b (3), a - xi b (4), xi a (5)
  template <int MODE>
  static __device__ __noinline__ void f2op(int d, int a, int b) {
    const int hh = half();
    Fq x, y, w;
    if (MODE <= 2) {
      ld(x, a + hh);
      if (MODE == 2) fq_dbl(x, x);
      else { ld(y, b + hh); if (MODE == 0) fq_add(x, x, y); else fq_sub(x, x, y); }
    } else {
      const int src = MODE == 5 ? a : b;
      ld(y, src + hh); ld(w, src + 1 - hh);
      xi_part(y, y, w, hh != 0);
      if (MODE == 5) x = y;
      else { ld(x, a + hh); if (MODE == 3) fq_add(x, x, y); else fq_sub(x, x, y); }
    }
    put(d + hh, x);
  }
  static __device__ __forceinline__ void f2add(int d, int a, int b) { f2op<0>(d, a, b); }
  static __device__ __forceinline__ void f2dbl(int d, int a) { f2op<2>(d, a, a); }
  static __device__ __forceinline__ void f2addxi(int d, int a, int b) { f2op<3>(d, a, b); }
  static __device__ __forceinline__ void f2mulxi(int d, int a) { f2op<5>(d, a, a); }
  // d = a - b - c (XI: a - b - xi c)
  template <bool XI>
  static __device__ __noinline__ void f2sub2(int d, int a, int b, int c) {
    const int hh = half();
    Fq x, y, z, w;
    ld(x, a + hh); ld(y, b + hh); ld(z, c + hh);
    if (XI) { ld(w, c + 1 - hh); xi_part(z, z, w, hh != 0); }
    fq_sub(x, x, y);
    fq_sub(x, x, z);
    put(d + hh, x);
  }

  // ---- two F_q^6 products at once: lane 0: d0 = a0 b0, lane 1: d1 = a1 b1 (f6mul of pairing_f_slots.cuh; x0 / x1:
  // two slots each holding xi (b_2); xi (b_1), used once, is made in registers).  No d may overlap any a, b or x.
  static __device__ __noinline__ void f6pair(int d0, int a0, int b0, int x0s, int d1, int a1, int b1, int x1s) {
    const bool hh = half() != 0;
    const int d = hh ? d1 : d0, a = hh ? a1 : a0, b = hh ? b1 : b0, xs = hh ? x1s : x0s;
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
      F2Acc g;
      Fq x0, x1, y0, y1;
      ld(x0, a); ld(x1, a + 1); ld(y0, b + 2 * k); ld(y1, b + 2 * k + 1);      // a_0 b_k
      f2a_term<0>(g, x0, x1, y0, y1);
      const int s1 = k >= 1 ? b + 2 * (k - 1) : xs;                            // a_1 b_(k-1)  or  a_1 (xi b_2)
      ld(x0, a + 2); ld(x1, a + 3); ld(y0, s1); ld(y1, s1 + 1);
      f2a_term<1>(g, x0, x1, y0, y1);
      const int s2 = k == 2 ? b : (k == 1 ? xs : b + 2);                       // a_2 b_0,  a_2 (xi b_2),  a_2 xi (b_1)
      ld(x0, a + 4); ld(x1, a + 5); ld(y0, s2); ld(y1, s2 + 1);
      if (k == 0) f2r_mul_xi(y0, y1);
      f2a_term<2>(g, x0, x1, y0, y1);
      f2a_finish<true>(x0, x1, g, c_f.qsqm[2]);
      st(d + 2 * k, x0); st(d + 2 * k + 1, x1);
    }
    sync();
  }
  // o = v * (c + L3 x^3 + L4 x^4): lane h makes coefficients 3h .. 3h+2 (line_mul of pairing_f_slots.cuh); o, v distinct
  static __device__ __noinline__ void line_mul(int o, int v) {
    const int k0 = 3 * half();
#pragma unroll 1
    for (int kk = 0; kk < 3; kk++) {
      const int k = k0 + kk;
      const int i3 = k >= 3 ? k - 3 : k + 3, i4 = k >= 4 ? k - 4 : k + 2;
      F2Acc g;
      Fq x0, x1, y0, y1;
      const int m3 = k >= 3 ? fpL3 : fpXL3, m4 = k >= 4 ? fpL4 : fpXL4;
      ld(x0, m3); ld(x1, m3 + 1);
      ld(y0, v + 2 * f12_pos(i3)); ld(y1, v + 2 * f12_pos(i3) + 1);
      f2a_term<0>(g, x0, x1, y0, y1);
      ld(x0, m4); ld(x1, m4 + 1);
      ld(y0, v + 2 * f12_pos(i4)); ld(y1, v + 2 * f12_pos(i4) + 1);
      f2a_term<1>(g, x0, x1, y0, y1);
      ld(x0, fpC);
      ld(y0, v + 2 * f12_pos(k)); ld(y1, v + 2 * f12_pos(k) + 1);
      fq_add_nr(x1, y0, y1);
      fqa_mac<false>(g.A, x0, y0);
      fqa_mac<false>(g.C, x0, x1);
      f2a_finish<true>(x0, x1, g, c_f.qsqm[1]);
      st(o + 2 * f12_pos(k), x0); st(o + 2 * f12_pos(k) + 1, x1);
    }
    sync();
  }
  // v <- v^2 in place (f12sqr of pairing_f_slots.cuh) with the 12-slot area t and 16 scratch slots at sc:
  //   t0 = A B -> t[0..5] (lane 0),  t1 = (A + B)(A + y B) -> sc[6..11] (lane 1);  A' = t1 - t0 - y t0,  B' = 2 t0
  static __device__ __forceinline__ void f12sqr(int v, int t, int sc) {
    f2add(t + 6, v, v + 6); f2add(t + 8, v + 2, v + 8); f2add(t + 10, v + 4, v + 10);        // A + B
    f2addxi(sc, v, v + 10); f2add(sc + 2, v + 2, v + 6); f2add(sc + 4, v + 4, v + 8);        // A + y B
    f2mulxi(sc + 12, v + 10);                                                                // xi B_2
    f2mulxi(sc + 14, sc + 4);                                                                // xi (A + y B)_2
    f6pair(t, v, v + 6, sc + 12, sc + 6, t + 6, sc, sc + 14);
    f2sub2<true>(v, sc + 6, t, t + 4);                         // A'0 = t1_0 - t0_0 - xi t0_2
    f2sub2<false>(v + 2, sc + 8, t + 2, t);                    // A'1 = t1_1 - t0_1 - t0_0
    f2sub2<false>(v + 4, sc + 10, t + 4, t + 2);               // A'2 = t1_2 - t0_2 - t0_1
    f2dbl(v + 6, t); f2dbl(v + 8, t + 2); f2dbl(v + 10, t + 4);
  }
};

// ---------------------------------------------------------------------------------------------
// k_f_prep: decode, validate, move to the basis in use (the prologue of k_f_miller_s), one thread per pairing.
// gq: [6 * kNS][n] words (Qx, Qy untwisted and scaled, P; Montgomery form, internal basis); flag[idx] = inputs usable.
// ---------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_f_prep(const uint8_t* __restrict__ P, const uint8_t* __restrict__ Q, uint32_t* __restrict__ flag,
         uint32_t* __restrict__ gq, size_t n, size_t stride1, const uint32_t* __restrict__ tab, size_t rows) {
  size_t idx = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (idx >= n) return;
  bool ok;
  Fq xP, yP, sg;
  if (tab) {
    ok = tab[3 * rows * kNS] != 0;             // fixed first argument: lines from the table (k_cc_pp_init)
    fq_zero(xP); fq_zero(yP);
  } else {
    const uint8_t* p = P + idx * stride1;
    fq_from_wire(xP, p);
    fq_from_wire(yP, p + kWS);
    ok = cc_on_curve(xP, yP);
  }
  F2 Qx, Qy, t, u;
  const uint8_t* q = Q + idx * (4 * kWS);
  fq_from_wire(Qx.a, q);
  fq_from_wire(Qx.b, q + kWS);
  fq_from_wire(Qy.a, q + 2 * kWS);
  fq_from_wire(Qy.b, q + 3 * kWS);
  fq_set(sg, c_f.sigma);
  fq_mul(Qx.b, Qx.b, sg);
  fq_mul(Qy.b, Qy.b, sg);
  f2_sqr(&t, &Qx);
  f2_mul(&t, &t, &Qx);
  f2_add(t, t, *f2_const(c_f.twist_b));
  f2_sqr(&u, &Qy);
  ok = ok && f2_eq(t, u);
  f2_mul(&Qx, &Qx, f2_const(c_f.kx));
  f2_mul(&Qy, &Qy, f2_const(c_f.ky));
  fq_st_global(gq, 0, n, idx, Qx.a); fq_st_global(gq, 1, n, idx, Qx.b);
  fq_st_global(gq, 2, n, idx, Qy.a); fq_st_global(gq, 3, n, idx, Qy.b);
  fq_st_global(gq, fgPx, n, idx, xP); fq_st_global(gq, fgPy, n, idx, yP);
  flag[idx] = ok ? 1u : 0u;
}

// The Miller loop proper: block = 2 BP threads, lanes (2j, 2j+1) share pairing blockIdx.x * BP + j.
template <int BP>
__global__ void __launch_bounds__(2 * BP, 256 / BP)
k_f_miller_p(uint32_t* __restrict__ mv, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ gq, size_t n,
             const uint32_t* __restrict__ tab, size_t rows) {
  using S = FP<BP>;
  size_t idx = (size_t)blockIdx.x * BP + (threadIdx.x >> 1);
  const bool live = idx < n;               // padding pairs of the last block run too (block-wide barriers below)
  if (!PBC_FS_LOCKSTEP && !live) return;
  if (!live) idx = 0;
  const uint32_t* g = gq + idx;
  S::qldg(fsX, g, fgPx, n, false);
  S::qldg(fsY, g, fgPy, n, false);
  S::qldc(fsZ, c_fp.one);
  S::qldc(fsV, c_fp.one);
#pragma unroll 1
  for (int s = 1; s < 12; s++) S::qzero(fsV + s);
  int V = fsV, T = fsT;
  size_t row = 0;
#if PBC_CC_NAF
  int m = (int)c_ccnaf.len - 2;
#else
  int m = (int)c_cc.rbits - 2;
#endif
  for (;;) {
    if (PBC_FS_LOCKSTEP) __syncthreads();
    if (tab) {
      // ---- fixed first argument: (a, b, c) of the next line from the table ----
      S::qldc(T + 4, tab + (3 * row + 0) * kNS); S::f2scale_g(fpL4, g, 0, n, T + 4);
      S::qldc(T + 4, tab + (3 * row + 1) * kNS); S::f2scale_g(fpL3, g, 2, n, T + 4);
      S::qldc(fpC, tab + (3 * row + 2) * kNS);
      S::f2mulxi(fpXL3, fpL3); S::f2mulxi(fpXL4, fpL4);
      row++;
    } else {
      // ---- tangent at V (a = -M Z^2, b = 2 Y Z^3, c = M X - 2 Y^2; the curve has A = 0), V <- 2V ----
      // the 14 products go out in pairs, one per lane (the update of V is also computed in the last
      // iteration, where nothing reads it)
      S::qmulh(T, fsZ, fsZ, T + 5, fsX, fsX);                             // Z^2, X^2
      S::qmulh(T + 2, fsY, fsY, T + 3, fsY, fsZ);                         // Y^2, Y Z
      S::qdbl(T + 1, T + 5); S::qadd(T + 1, T + 1, T + 5);                // M = 3 X^2
      S::qmulh(T + 4, T + 1, T, fpC, T + 1, fsX);                         // M Z^2, M X
      S::qneg(T + 4, T + 4);                                              // a
      S::f2scale_g(fpL4, g, 0, n, T + 4);                                 // L4 = Qx a
      S::qdbl(T + 3, T + 3);                                              // Z' = 2 Y Z
      S::qmulh(T + 4, T + 3, T, T + 5, fsX, T + 2);                       // b = Z' Z^2, X Y^2
      S::f2scale_g(fpL3, g, 2, n, T + 4);                                 // L3 = Qy b
      S::qsub(fpC, fpC, T + 2); S::qsub(fpC, fpC, T + 2);                 // c = M X - 2 Y^2
      S::f2mulxi(fpXL3, fpL3); S::f2mulxi(fpXL4, fpL4);
      S::qcopy(fsZ, T + 3);
      S::qmulh(fsX, T + 1, T + 1, T + 2, T + 2, T + 2);                   // M^2, Y^4
      S::qdbl(T + 5, T + 5, 2);                                           // S = 4 X Y^2
      S::qsub(fsX, fsX, T + 5); S::qsub(fsX, fsX, T + 5);                 // X' = M^2 - 2 S
      S::qdbl(T + 2, T + 2, 3);                                           // 8 Y^4
      S::qsub(T + 5, T + 5, fsX);
      S::qmulh(fsY, T + 1, T + 5, fsY, T + 1, T + 5);                     // M (S - X')
      S::qsub(fsY, fsY, T + 2);                                           // Y'
    }
    S::line_mul(T, V);
    { int s = V; V = T; T = s; }
    if (m == 0) break;
#if PBC_CC_NAF
    if ((c_ccnaf.nz[m >> 5] >> (m & 31)) & 1u) {
      const bool minus = (c_ccnaf.neg[m >> 5] >> (m & 31)) & 1u;
#else
    if ((c_cc.r[m >> 5] >> (m & 31)) & 1u) {
      const bool minus = false;
#endif
      if (tab) {
        S::qldc(T + 4, tab + (3 * row + 0) * kNS); S::f2scale_g(fpL4, g, 0, n, T + 4);
        S::qldc(T + 4, tab + (3 * row + 1) * kNS); S::f2scale_g(fpL3, g, 2, n, T + 4);
        S::qldc(fpC, tab + (3 * row + 2) * kNS);
        S::f2mulxi(fpXL3, fpL3); S::f2mulxi(fpXL4, fpL4);
        row++;
      } else {
        // ---- chord through V and +-P (a = Y - yS Z^3, b = (xP Z^2 - X) Z, c = yS Z X - xP Y), V <- V +- P ----
        S::qldg(T + 6, g, fgPx, n, false);                 // xP
        S::qldg(T + 7, g, fgPy, n, minus);                 // yS
        S::qmulh(T, fsZ, fsZ, T + 8, T + 7, fsZ);                          // Z^2, yS Z
        S::qmulh(T + 9, T + 6, fsY, T + 1, T, fsZ);                        // xP Y, Z^3
        S::qmulh(T + 2, T + 6, T, T + 8, T + 8, fsX);                      // xP Z^2, yS Z X
        S::qsub(T + 2, T + 2, fsX);                                        // H = xP Z^2 - X
        S::qmulh(T + 3, T + 7, T + 1, T + 4, T + 2, fsZ);                  // yS Z^3, b = H Z
        S::qsub(T + 5, fsY, T + 3);                                        // a = Y - yS Z^3
        S::f2scale_g(fpL4, g, 0, n, T + 5);
        S::qsub(T + 3, T + 3, fsY);                                        // R = yS Z^3 - Y
        S::f2scale_g(fpL3, g, 2, n, T + 4);
        S::qsub(fpC, T + 8, T + 9);                                        // c
        S::f2mulxi(fpXL3, fpL3); S::f2mulxi(fpXL4, fpL4);
        S::qcopy(fsZ, T + 4);                                              // Z of the sum
        S::qmulh(T, T + 2, T + 2, T + 10, T + 3, T + 3);                   // H^2, R^2
        S::qmulh(T + 1, T, T + 2, T, T, fsX);                              // H^3, X H^2
        S::qsub(fsX, T + 10, T + 1); S::qsub(fsX, fsX, T); S::qsub(fsX, fsX, T);   // X3 = R^2 - H^3 - 2 X H^2
        S::qsub(T, T, fsX);
        S::qmulh(T, T, T + 3, T + 1, T + 1, fsY);                          // R (X H^2 - X3), H^3 Y
        S::qsub(fsY, T, T + 1);                                            // Y3
      }
      S::line_mul(T, V);
      { int s = V; V = T; T = s; }
    }
    m--;
    S::f12sqr(V, T, fpScratch);
  }
  if (!live) return;
  // publish: lane h stores the slots of its parity (flagged-off inputs: the identity)
  const bool ok = flag[idx] != 0;
  Fq x;
#pragma unroll 1
  for (int s = S::half(); s < 12; s += 2) {
    S::ld(x, V + s);
    if (!ok) { if (s == 0) fq_one(x); else fq_zero(x); }
    fq_st_global(mv, s, n, idx, x);
  }
}

}  // namespace pbcb200
