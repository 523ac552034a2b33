// ubench.cu -- instruction-rate probes for the integer pipe of sm_100a (design aid; results in
// profiles/).  Each kernel runs `iters` x 4 x 8 independent operations per thread.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define BODY(stmt)                                   \
  for (int i = 0; i < iters; i++) {                  \
    _Pragma("unroll") for (int u = 0; u < 4; u++) {  \
      _Pragma("unroll") for (int k = 0; k < 8; k++) { stmt; } } }

// plain 64-bit accumulate, no carries
__global__ void k_wide(uint32_t* out, uint32_t s, int iters) {
  uint32_t x = s + threadIdx.x, y = s * 3 + blockIdx.x;
  uint64_t acc[8]; for (int k = 0; k < 8; k++) acc[k] = k;
  BODY(asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"((uint32_t)acc[(k + 1) & 7] ^ x), "r"(y)))
  uint64_t r = 0; for (int k = 0; k < 8; k++) r ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)r ^ (uint32_t)(r >> 32);
}
// carry-out only: IMAD.WIDE R, P + IADD3.X absorbing (product-scanning MAC)
__global__ void k_wide_cout(uint32_t* out, uint32_t s, int iters) {
  uint32_t x = s + threadIdx.x, y = s * 3 + blockIdx.x;
  uint32_t l[8], h[8], t[8]; for (int k = 0; k < 8; k++) { l[k] = k; h[k] = k; t[k] = 0; }
  BODY(asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                    : "+r"(l[k]), "+r"(h[k]), "+r"(t[k]) : "r"(l[(k + 1) & 7]), "r"(y)))
  uint32_t r = 0; for (int k = 0; k < 8; k++) r ^= l[k] ^ h[k] ^ t[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// carry chains of length 4 pairs (operand-scanning style): IMAD.WIDE.X with P in and out
__global__ void k_wide_chain(uint32_t* out, uint32_t s, int iters) {
  uint32_t x = s + threadIdx.x, y = s * 3 + blockIdx.x;
  uint32_t a[16]; for (int k = 0; k < 16; k++) a[k] = k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[0]), "+r"(a[1]) : "r"(a[15]), "r"(y));
#pragma unroll
      for (int k = 2; k < 16; k += 2)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[k]), "+r"(a[k + 1]) : "r"(a[k - 1]), "r"(y));
    }
  }
  uint32_t r = 0; for (int k = 0; k < 16; k++) r ^= a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// 32-bit IMAD (lo) only
__global__ void k_imad_lo(uint32_t* out, uint32_t s, int iters) {
  uint32_t x = s + threadIdx.x, y = s * 3 + blockIdx.x;
  uint32_t acc[8]; for (int k = 0; k < 8; k++) acc[k] = k;
  BODY(asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[(k + 1) & 7]), "r"(y)))
  uint32_t r = 0; for (int k = 0; k < 8; k++) r ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// IMAD.HI only
__global__ void k_imad_hi(uint32_t* out, uint32_t s, int iters) {
  uint32_t x = s + threadIdx.x, y = s * 3 + blockIdx.x;
  uint32_t acc[8]; for (int k = 0; k < 8; k++) acc[k] = k;
  BODY(asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[(k + 1) & 7] | 0x80000000u), "r"(y | 0x80000000u)))
  uint32_t r = 0; for (int k = 0; k < 8; k++) r ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// plain IMAD.WIDE interleaved 1:1 with independent IADD3 (dual-pipe issue test)
__global__ void k_wide_plus_add(uint32_t* out, uint32_t s, int iters) {
  uint32_t x = s + threadIdx.x, y = s * 3 + blockIdx.x;
  uint64_t acc[8]; uint32_t b[8]; for (int k = 0; k < 8; k++) { acc[k] = k; b[k] = k; }
  BODY(asm volatile("mad.wide.u32 %0, %2, %3, %0; add.u32 %1, %1, %2;" : "+l"(acc[k]), "+r"(b[k]) : "r"((uint32_t)acc[(k + 1) & 7] ^ x), "r"(y)))
  uint64_t r = 0; for (int k = 0; k < 8; k++) r ^= acc[k] ^ b[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)r ^ (uint32_t)(r >> 32);
}
// plain IMAD.WIDE interleaved 1:2 with IADD3 / shift work
__global__ void k_wide_plus_2add(uint32_t* out, uint32_t s, int iters) {
  uint32_t x = s + threadIdx.x, y = s * 3 + blockIdx.x;
  uint64_t acc[8]; uint32_t b[8], c[8]; for (int k = 0; k < 8; k++) { acc[k] = k; b[k] = k; c[k] = k; }
  BODY(asm volatile("mad.wide.u32 %0, %3, %4, %0; add.u32 %1, %1, %3; xor.b32 %2, %2, %1;" : "+l"(acc[k]), "+r"(b[k]), "+r"(c[k]) : "r"((uint32_t)acc[(k + 1) & 7] ^ x), "r"(y)))
  uint64_t r = 0; for (int k = 0; k < 8; k++) r ^= acc[k] ^ b[k] ^ c[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)r ^ (uint32_t)(r >> 32);
}
// 64-bit add with carry through predicate only (IADD3 + IADD3.X pairs)
__global__ void k_add64(uint32_t* out, uint32_t s, int iters) {
  uint32_t x = s + threadIdx.x, y = s * 3 + blockIdx.x;
  uint64_t acc[8]; for (int k = 0; k < 8; k++) acc[k] = k;
  BODY(acc[k] += acc[(k + 1) & 7] ^ y)
  uint64_t r = 0; for (int k = 0; k < 8; k++) r ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)r ^ (uint32_t)(r >> 32);
}

// FP64 FMA rate (context: B200 keeps a full-rate FP64 pipe)
__global__ void k_dfma(uint32_t* out, uint32_t s, int iters) {
  double x = 1.0 + 1e-9 * (s + threadIdx.x), y = 1.0 - 1e-9 * (s * 3 + blockIdx.x);
  double acc[8]; for (int k = 0; k < 8; k++) acc[k] = k;
  BODY(acc[k] = fma(acc[(k + 1) & 7], y, x))
  double r = 0; for (int k = 0; k < 8; k++) r += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double2ll_rn(r);
}
// FP32 FMA rate
__global__ void k_ffma(uint32_t* out, uint32_t s, int iters) {
  float x = 1.0f + 1e-6f * (s + threadIdx.x), y = 1.0f - 1e-6f * (s * 3 + blockIdx.x);
  float acc[8]; for (int k = 0; k < 8; k++) acc[k] = k;
  BODY(acc[k] = fmaf(acc[(k + 1) & 7], y, x))
  float r = 0; for (int k = 0; k < 8; k++) r += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)r;
}
template <class K> static void run(const char* name, K kern, double ops_per_iter, int sms) {
  int blocks = sms * 8, threads = 256, iters = 2000;
  uint32_t* out; cudaMalloc(&out, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(out, 7, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 3; r++) kern<<<blocks, threads>>>(out, 7 + r, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  double ops = (double)blocks * threads * iters * ops_per_iter;
  printf("{\"ubench\": \"%s\", \"ms\": %.4f, \"ops_per_s\": %.4e, \"per_clk_per_sm_at_1.9GHz\": %.2f}\n", name, ms,
         ops / (ms * 1e-3), ops / (ms * 1e-3) / sms / 1.9e9);
  cudaFree(out);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  run("imad_wide_plain", k_wide, 32, sms);
  run("imad_wide_carry_out+iadd3x (products)", k_wide_cout, 32, sms);
  run("imad_wide_x chain (products)", k_wide_chain, 32, sms);
  run("imad_lo32", k_imad_lo, 32, sms);
  run("imad_hi32", k_imad_hi, 32, sms);
  run("imad_wide_plain + 1 iadd (wide count)", k_wide_plus_add, 32, sms);
  run("imad_wide_plain + 2 alu (wide count)", k_wide_plus_2add, 32, sms);
  run("add64 (iadd3+iadd3.x pairs)", k_add64, 32, sms);
  run("dfma", k_dfma, 32, sms);
  run("ffma", k_ffma, 32, sms);
  return 0;
}
