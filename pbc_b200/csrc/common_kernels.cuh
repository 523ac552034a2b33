// common_kernels.cuh -- batch inversion (replaces the per-pairing mpz_invert calls of
// arith/montfp.c:401-422) and the F_p / IMAD micro-kernels used by bench.py's roofline.
#pragma once
#include "slots.cuh"

namespace pbcb200 {

// r = a^(p-2) on slots (Fermat; the modulus is prime: arith/fp.c:36-49 requires it too).
template <class O, int N>
__device__ __forceinline__ void slot_fermat_inverse(int r, int a, int one_slot) {
  O::copy(r, one_slot);
  int top = N * 32 - 1;
  while (top > 0 && !((c_fp.pm2[top >> 5] >> (top & 31)) & 1u)) top--;
  for (int j = top; j >= 0; j--) {
    O::sqr(r, r);
    if ((c_fp.pm2[j >> 5] >> (j & 31)) & 1u) O::mul(r, r, a);
  }
}

// In-place inversion of n Montgomery-form elements d[VPE][n] (limb-major vectors).
// Montgomery's simultaneous-inversion trick (the reference uses it only inside
// element_multi_double, ecc/curve.c:210-281): thread t owns the chain {t, t+T, t+2T, ...},
// T = number of chains, so every access is coalesced; one Fermat inversion per chain.
// Zero elements stay zero and do not poison their chain.  prefix: scratch of the same shape.
template <int N, bool FULL, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_batch_invert(void* __restrict__ d, void* __restrict__ prefix, size_t n, size_t T) {
  using O = Ops<N, FULL, BLOCK>;
  enum { sACC, sE, sINV, sONE, sT, kSlots };
  size_t t = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (t >= T) return;
  O::set_const(sONE, c_fp.one);
  O::copy(sACC, sONE);
  size_t K = (n + T - 1) / T;
  for (size_t j = 0; j < K; j++) {
    size_t idx = j * T + t;
    if (idx >= n) break;
    O::ld_global(sE, d, 0, n, idx);
    if (!O::is_zero(sE)) O::mul(sACC, sACC, sE);
    O::st_global(prefix, 0, n, idx, sACC);
  }
  slot_fermat_inverse<O, N>(sINV, sACC, sONE);
  for (size_t j = K; j-- > 0;) {
    size_t idx = j * T + t;
    if (idx >= n) continue;
    O::ld_global(sE, d, 0, n, idx);
    if (O::is_zero(sE)) continue;                 // stays zero; chain product skipped it
    if (j > 0) O::ld_global(sT, prefix, 0, n, idx - T); else O::copy(sT, sONE);
    O::mul(sT, sT, sINV);                         // 1/e_j = inv * prefix_{j-1}
    O::mul(sINV, sINV, sE);
    O::st_global(d, 0, n, idx, sT);
  }
}

// ---- micro-kernels for the integer-pipe roofline --------------------------------------------
// Dependent chain of `iters` Montgomery multiplications per thread, operands in registers.
template <int N, bool FULL>
__global__ void k_fpmul_chain(uint32_t* __restrict__ out, const uint32_t* __restrict__ in,
                              int iters) {
  size_t T = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t a[N], b[N];
#pragma unroll
  for (int k = 0; k < N; k++) { a[k] = in[(2 * k) * T + t]; b[k] = in[(2 * k + 1) * T + t]; }
  for (int i = 0; i < iters; i++) mont_mul<N, FULL>(a, a, b);
#pragma unroll
  for (int k = 0; k < N; k++) out[k * T + t] = a[k];
}

// Same chain through the shared-memory slot machine (what the pairing kernels execute).
template <int N, bool FULL, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
k_fpmul_slots(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, int iters) {
  using O = Ops<N, FULL, BLOCK>;
  size_t T = (size_t)gridDim.x * BLOCK, t = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  uint32_t a[N], b[N];
#pragma unroll
  for (int k = 0; k < N; k++) { a[k] = in[(2 * k) * T + t]; b[k] = in[(2 * k + 1) * T + t]; }
  O::st(0, a); O::st(1, b);
  for (int i = 0; i < iters; i++) O::mul(0, 0, 1);
  O::ld(a, 0);
#pragma unroll
  for (int k = 0; k < N; k++) out[k * T + t] = a[k];
}

// Dependency-free IMAD.WIDE.U32 stream: 8 independent 64-bit accumulators per thread.
// Measures the issue-rate ceiling that bounds every kernel above ("measure peak
// IMAD/s with a dependency-free microkernel").
__global__ void k_imad_peak(uint64_t* __restrict__ out, uint32_t seed, int iters) {
  uint32_t x = seed + threadIdx.x, y = seed * 3u + blockIdx.x;
  uint64_t acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x + k), "r"(y));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= acc[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace pbcb200
