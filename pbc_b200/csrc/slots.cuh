// slots.cuh -- per-thread "field register file" in shared memory + out-of-line field operations.
//
// One pairing runs per thread.  A 512-bit F_q element is 16 registers; the Miller loop of
// ecc/a_param.c:1053-1198 keeps ~14 of them live, which does not fit the register file at any
// useful occupancy and, fully inlined, would unroll to hundreds of KB of SASS.  Instead every
// long-lived element lives in a *slot* of shared memory and each field operation is a single
// out-of-line routine (one copy of the ~560-IMAD multiplier in the instruction cache):
//
//     slot s, 16-byte vector v, thread t  ->  smem[(s * VPE + v) * BLOCK + t]   (uint4 / uint2)
//
// so a warp's access to one vector is 32 consecutive 16-byte words: conflict-free LDS.128 /
// STS.128.  An operation loads its operands (8 LDS.128), runs entirely in registers, and stores
// the result (4 STS.128): ~12 shared-memory instructions against ~540 IMAD.WIDE.
#pragma once
#include "fp.cuh"

namespace pbcb200 {

extern __shared__ uint4 pbc_smem[];

template <int N, bool FULL, int BLOCK>
struct Ops {
  static constexpr int kVecWords = (N % 4 == 0) ? 4 : 2;
  static constexpr int kVecs = N / kVecWords;
  static constexpr int kSlotBytes = N * 4 * BLOCK;

  static __device__ __forceinline__ void ld(uint32_t* r, int s) {
    if constexpr (kVecWords == 4) {
      const uint4* b = pbc_smem + s * (kVecs * BLOCK) + threadIdx.x;
#pragma unroll
      for (int v = 0; v < kVecs; v++) {
        uint4 q = b[v * BLOCK];
        r[4 * v] = q.x; r[4 * v + 1] = q.y; r[4 * v + 2] = q.z; r[4 * v + 3] = q.w;
      }
    } else {
      const uint2* b = reinterpret_cast<const uint2*>(pbc_smem) + s * (kVecs * BLOCK) + threadIdx.x;
#pragma unroll
      for (int v = 0; v < kVecs; v++) {
        uint2 q = b[v * BLOCK];
        r[2 * v] = q.x; r[2 * v + 1] = q.y;
      }
    }
  }
  static __device__ __forceinline__ void st(int s, const uint32_t* r) {
    if constexpr (kVecWords == 4) {
      uint4* b = pbc_smem + s * (kVecs * BLOCK) + threadIdx.x;
#pragma unroll
      for (int v = 0; v < kVecs; v++)
        b[v * BLOCK] = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
    } else {
      uint2* b = reinterpret_cast<uint2*>(pbc_smem) + s * (kVecs * BLOCK) + threadIdx.x;
#pragma unroll
      for (int v = 0; v < kVecs; v++) b[v * BLOCK] = make_uint2(r[2 * v], r[2 * v + 1]);
    }
  }

  // ---- out-of-line operations on slots (d may alias a or b) ----
  static __device__ __noinline__ void mul(int d, int a, int b) {
    uint32_t x[N], y[N];
    ld(x, a); ld(y, b);
    fp_mul<N, FULL>(x, x, y);
    st(d, x);
  }
  static __device__ __noinline__ void sqr(int d, int a) {
    uint32_t x[N];
    ld(x, a);
    fp_sqr<N, FULL>(x, x);
    st(d, x);
  }
  // ---- ONE multiplier routine for the kernels that need operand variants (k_a_miller9) ----
  // d = x * y with  x = a [+ a2],  y = b [+ b2 | - b2]  or  y = the element at g (limb-major global
  // array, vectors n apart: a value used once or twice per loop iteration lives in L2 instead of a
  // slot).  All branches are warp-uniform.  One copy of the 528-product body instead of one per
  // variant: with 12 warps per SM the four separate copies (mul, mul2<+>, mul2<->, mulg: 3 200
  // instructions, 51 KB) made instruction fetch the second largest stall (ncu, round 2).
  static constexpr uint32_t MX_ADD_A = 1u, MX_ADD_B = 2u, MX_SUB_B = 4u, MX_GLOBAL_B = 8u;
  static __device__ __noinline__ void mulx(int d, int a, int b, uint32_t mode, int a2, int b2, const void* g, size_t n) {
    uint32_t x[N], y[N];
    if (mode & MX_GLOBAL_B) {
      if constexpr (kVecWords == 4) {
        const uint4* p = reinterpret_cast<const uint4*>(g);
#pragma unroll
        for (int v = 0; v < kVecs; v++) {
          uint4 q = p[v * n];
          y[4 * v] = q.x; y[4 * v + 1] = q.y; y[4 * v + 2] = q.z; y[4 * v + 3] = q.w;
        }
      } else {
        const uint2* p = reinterpret_cast<const uint2*>(g);
#pragma unroll
        for (int v = 0; v < kVecs; v++) {
          uint2 q = p[v * n];
          y[2 * v] = q.x; y[2 * v + 1] = q.y;
        }
      }
    } else {
      ld(y, b);
    }
    if (mode & (MX_ADD_B | MX_SUB_B)) {
      ld(x, b2);
      if (mode & MX_ADD_B) fp_add<N, FULL>(y, y, x); else fp_sub<N>(y, y, x);
    }
    ld(x, a);
    if (mode & MX_ADD_A) {
      uint32_t z[N];
      ld(z, a2);
      fp_add<N, FULL>(x, x, z);
    }
    fp_mul<N, FULL>(x, x, y);
    st(d, x);
  }
  template <bool SUBB>
  static __device__ __forceinline__ void mul2(int d, int a, int a2, int b, int b2) {
    mulx(d, a, b, MX_ADD_A | (SUBB ? MX_SUB_B : MX_ADD_B), a2, b2, nullptr, 0);
  }
  static __device__ __forceinline__ void mulg(int d, int a, const void* g, size_t n) {
    mulx(d, a, 0, MX_GLOBAL_B, 0, 0, g, n);
  }
  static __device__ __forceinline__ void mulp(int d, int a, int b) { mulx(d, a, b, 0, 0, 0, nullptr, 0); }
  // d = a*b - c
  static __device__ __noinline__ void mulsub(int d, int a, int b, int c) {
    uint32_t x[N], y[N];
    ld(x, a); ld(y, b);
    fp_mul<N, FULL>(x, x, y);
    ld(y, c);
    fp_sub<N>(x, x, y);
    st(d, x);
  }
  // d = a^2 - c
  static __device__ __noinline__ void sqrsub(int d, int a, int c) {
    uint32_t x[N], y[N];
    ld(x, a);
    fp_sqr<N, FULL>(x, x);
    ld(y, c);
    fp_sub<N>(x, x, y);
    st(d, x);
  }
  // ---- fused multiply: d = ((a [+- a2]) * (b [+- b2]) [+- n1 c1] [+- n2 c2]) * 2^k ----
  // One call replaces a multiplication and the additive operations around it (each of which
  // would otherwise be its own call with its own shared-memory round trip).  `f` is built with
  // the F_* helpers below; all branches on it are warp-uniform.
  static constexpr uint32_t F_ADD_A = 1u, F_SUB_A = 2u, F_ADD_B = 4u, F_SUB_B = 8u;
  static __host__ __device__ constexpr uint32_t F_ADD_C1(uint32_t n) { return n << 4; }
  static __host__ __device__ constexpr uint32_t F_SUB_C1(uint32_t n) { return (n << 4) | 0x80u; }
  static __host__ __device__ constexpr uint32_t F_ADD_C2(uint32_t n) { return n << 8; }
  static __host__ __device__ constexpr uint32_t F_SUB_C2(uint32_t n) { return (n << 8) | 0x800u; }
  static __host__ __device__ constexpr uint32_t F_DBL(uint32_t k) { return k << 12; }

  static __device__ __forceinline__ void post_ops(uint32_t* x, uint32_t f, int c1, int c2) {
    uint32_t y[N];
    if (f & 0x70u) {
      ld(y, c1);
      for (uint32_t i = (f >> 4) & 7u; i; i--) {
        if (f & 0x80u) fp_sub<N>(x, x, y); else fp_add<N, FULL>(x, x, y);
      }
    }
    if (f & 0x700u) {
      ld(y, c2);
      for (uint32_t i = (f >> 8) & 7u; i; i--) {
        if (f & 0x800u) fp_sub<N>(x, x, y); else fp_add<N, FULL>(x, x, y);
      }
    }
    for (uint32_t i = (f >> 12) & 3u; i; i--) fp_add<N, FULL>(x, x, x);
  }
  static __device__ __noinline__ void fmul(int d, int a, int b, uint32_t f, int a2, int b2, int c1, int c2) {
    uint32_t x[N], y[N];
    ld(y, b);
    if (f & (F_ADD_B | F_SUB_B)) {
      ld(x, b2);
      if (f & F_ADD_B) fp_add<N, FULL>(y, y, x); else fp_sub<N>(y, y, x);
    }
    ld(x, a);
    if (f & (F_ADD_A | F_SUB_A)) {
      uint32_t z[N];
      ld(z, a2);
      if (f & F_ADD_A) fp_add<N, FULL>(x, x, z); else fp_sub<N>(x, x, z);
    }
    fp_mul<N, FULL>(x, x, y);
    post_ops(x, f, c1, c2);
    st(d, x);
  }
  static __device__ __noinline__ void fsqr(int d, int a, uint32_t f, int c1, int c2) {
    uint32_t x[N];
    ld(x, a);
    fp_sqr<N, FULL>(x, x);
    post_ops(x, f, c1, c2);
    st(d, x);
  }
  static __device__ __noinline__ void add(int d, int a, int b) {
    uint32_t x[N], y[N];
    ld(x, a); ld(y, b);
    fp_add<N, FULL>(x, x, y);
    st(d, x);
  }
  static __device__ __noinline__ void sub(int d, int a, int b) {
    uint32_t x[N], y[N];
    ld(x, a); ld(y, b);
    fp_sub<N>(x, x, y);
    st(d, x);
  }
  // d = 2^k * a
  static __device__ __noinline__ void dbl(int d, int a, int k = 1) {
    uint32_t x[N];
    ld(x, a);
    for (int i = 0; i < k; i++) fp_add<N, FULL>(x, x, x);
    st(d, x);
  }
  static __device__ __noinline__ void neg(int d, int a) {
    uint32_t x[N];
    ld(x, a);
    fp_neg<N>(x, x);
    st(d, x);
  }
  static __device__ __noinline__ void halve(int d, int a, int k = 1) {
    uint32_t x[N];
    ld(x, a);
    for (int i = 0; i < k; i++) fp_halve<N, FULL>(x, x);
    st(d, x);
  }
  static __device__ __forceinline__ void copy(int d, int a) {
    uint32_t x[N];
    ld(x, a);
    st(d, x);
  }
  static __device__ __forceinline__ void set_const(int d, const uint32_t* c) {
    uint32_t x[N];
#pragma unroll
    for (int k = 0; k < N; k++) x[k] = c[k];
    st(d, x);
  }
  static __device__ __forceinline__ bool is_zero(int a) {
    uint32_t x[N];
    ld(x, a);
    return fp_is_zero<N>(x);
  }
  static __device__ __forceinline__ bool eq(int a, int b) {
    uint32_t x[N], y[N];
    ld(x, a); ld(y, b);
    return fp_eq<N>(x, y);
  }

  // ---- global <-> slot, limb-major batch arrays:  g[(e * kVecs + v) * n + idx]  ----
  static __device__ __forceinline__ void ld_global(int s, const void* g, int e, size_t n, size_t idx) {
    uint32_t x[N];
    if constexpr (kVecWords == 4) {
      const uint4* b = reinterpret_cast<const uint4*>(g) + (size_t)e * kVecs * n + idx;
#pragma unroll
      for (int v = 0; v < kVecs; v++) {
        uint4 q = b[v * n];
        x[4 * v] = q.x; x[4 * v + 1] = q.y; x[4 * v + 2] = q.z; x[4 * v + 3] = q.w;
      }
    } else {
      const uint2* b = reinterpret_cast<const uint2*>(g) + (size_t)e * kVecs * n + idx;
#pragma unroll
      for (int v = 0; v < kVecs; v++) {
        uint2 q = b[v * n];
        x[2 * v] = q.x; x[2 * v + 1] = q.y;
      }
    }
    st(s, x);
  }
  static __device__ __forceinline__ void st_global(void* g, int e, size_t n, size_t idx, int s) {
    uint32_t x[N];
    ld(x, s);
    if constexpr (kVecWords == 4) {
      uint4* b = reinterpret_cast<uint4*>(g) + (size_t)e * kVecs * n + idx;
#pragma unroll
      for (int v = 0; v < kVecs; v++)
        b[v * n] = make_uint4(x[4 * v], x[4 * v + 1], x[4 * v + 2], x[4 * v + 3]);
    } else {
      uint2* b = reinterpret_cast<uint2*>(g) + (size_t)e * kVecs * n + idx;
#pragma unroll
      for (int v = 0; v < kVecs; v++) b[v * n] = make_uint2(x[2 * v], x[2 * v + 1]);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// wire format (arith/montfp.c:487-517): big-endian, fixed width WB bytes per F_q coordinate.
// ---------------------------------------------------------------------------------------------
// bytes -> little-endian 32-bit limbs (not yet reduced, not yet Montgomery)
template <int N, int WB>
__device__ __forceinline__ void limbs_from_be(uint32_t* x, const uint8_t* p) {
  static_assert(WB % 4 == 0 && WB <= 4 * N, "coordinate width");
#pragma unroll
  for (int k = 0; k < N; k++) {
    if (k < WB / 4) {
      uint32_t w = *reinterpret_cast<const uint32_t*>(p + WB - 4 - 4 * k);
      x[k] = __byte_perm(w, 0, 0x0123);
    } else {
      x[k] = 0;
    }
  }
}
// the same from a byte pointer of any alignment
template <int N, int WB>
__device__ __forceinline__ void limbs_from_be_bytes(uint32_t* x, const uint8_t* p) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    if (k < WB / 4) {
      const uint8_t* b = p + WB - 4 - 4 * k;
      x[k] = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    } else {
      x[k] = 0;
    }
  }
}
template <int N, int WB>
__device__ __forceinline__ void limbs_to_be(uint8_t* p, const uint32_t* x) {
#pragma unroll
  for (int k = 0; k < WB / 4; k++)
    *reinterpret_cast<uint32_t*>(p + WB - 4 - 4 * k) = __byte_perm(x[k], 0, 0x0123);
}

}  // namespace pbcb200
