// cuda_sim.hpp -- just enough of the CUDA runtime and device environment to run the product's
// kernels on the CPU, one simulated thread at a time.
//
// TEST INFRASTRUCTURE (tests/host/make_host_sim.py, tests/test_kernels_on_cpu_sim.py).  "Device"
// memory is host memory, every copy and launch is synchronous, a launch runs the kernel function
// once per (block, thread) with blockIdx / threadIdx set.  Threads of a block therefore never run
// concurrently: fine for these kernels (one pairing per thread; the only barriers are the lock-step
// ones of miller_cc.cuh, which order instruction fetch, not data).  Shared memory is one static
// buffer per host thread, sized like the largest opt-in allocation of an SM.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <condition_variable>
#include <mutex>
#include <thread>

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a) : x(a) {} dim3(int a) : x((unsigned)a) {} dim3(size_t a) : x((unsigned)a) {} };

namespace cusim {
// per host thread: the engine drives each simulated device from its own thread
inline dim3& tid() { static thread_local dim3 v; return v; }
inline dim3& bid() { static thread_local dim3 v; return v; }
inline dim3& bdim() { static thread_local dim3 v; return v; }
inline dim3& gdim() { static thread_local dim3 v; return v; }
inline int& current_device() { static thread_local int d = 0; return d; }
// PBC_SIM_DEVICES=N makes the simulator report N devices (they share the host's memory)
inline int device_count() { const char* e = getenv("PBC_SIM_DEVICES"); int n = e ? atoi(e) : 1; return n > 0 ? n : 1; }
inline uint64_t& launches() { static uint64_t n = 0; return n; }
// optional cap on simulated threads per launch (0 = all): kernels whose padding threads keep
// computing (k_a_miller's `live` pattern) would otherwise cost a full block per launch
inline size_t& live_threads() { static size_t n = 0; return n; }
// shared memory is poisoned at the start of every block and device allocations at birth, so a
// kernel that reads what it never wrote sees garbage here as it would on the GPU
inline void (*&poison_shared())() { static void (*f)() = nullptr; return f; }
// this host thread's own shared-memory buffer (set by the generated source, which owns the buffer)
inline void* (*&shared_self_fn())() { static void* (*f)() = nullptr; return f; }
inline void* shared_self() { return shared_self_fn() ? shared_self_fn()() : nullptr; }
// ---- lane pairs ----
// A kernel whose adjacent lanes (2j, 2j+1) cooperate through shared memory, __syncwarp() and
// __shfl_xor_sync(.., 1) (the pair kernels of pairing_f_pair.cuh) is launched with pair_mode() set: the two
// lanes of a pair then run on two host threads that meet at every __syncwarp / shuffle; pairs still run one
// after the other.  Both lanes must reach the same sequence of barriers (true of those kernels: one
// instruction stream, the lane parity only selects operands).
struct PairCtx {
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0;
  unsigned long long phase = 0;
  uint32_t box[2] = {0, 0};
  void sync() {
    std::unique_lock<std::mutex> lk(m);
    unsigned long long ph = phase;
    if (++waiting == 2) { waiting = 0; phase++; cv.notify_all(); }
    else cv.wait(lk, [&] { return phase != ph; });
  }
};
inline bool& pair_mode() { static thread_local bool v = false; return v; }
inline PairCtx*& pair_ctx() { static thread_local PairCtx* p = nullptr; return p; }
// the shared-memory buffer a simulated thread sees: its host thread's own, or (second lane of a pair) the first lane's
inline void*& shared_override() { static thread_local void* p = nullptr; return p; }

template <class F>
inline void launch(dim3 grid, dim3 block, F&& body) {
  launches()++;
  gdim() = grid; bdim() = block;
  size_t done = 0;
  const bool pairs = pair_mode();
  for (unsigned b = 0; b < grid.x; b++) {
    if (poison_shared()) poison_shared()();
    for (unsigned t = 0; t < block.x; t++) {
      if (live_threads() && done >= live_threads()) return;
      bid() = dim3(b); tid() = dim3(t);
      if (pairs && t + 1 < block.x) {
        PairCtx ctx;
        void* shared = shared_self();
        const dim3 g = grid, bd = block;
        std::thread other([&, b, t, g, bd, shared] {
          gdim() = g; bdim() = bd; bid() = dim3(b); tid() = dim3(t + 1);
          pair_ctx() = &ctx; shared_override() = shared;
          body();
          pair_ctx() = nullptr; shared_override() = nullptr;
        });
        pair_ctx() = &ctx;
        body();
        other.join();
        pair_ctx() = nullptr;
        t++; done += 2;
        continue;
      }
      body();
      done++;
    }
  }
}
}  // namespace cusim
#define threadIdx (::cusim::tid())
#define blockIdx (::cusim::bid())
#define blockDim (::cusim::bdim())
#define gridDim (::cusim::gdim())

static inline void __syncthreads() {}
static inline void __syncwarp() { if (::cusim::pair_ctx()) ::cusim::pair_ctx()->sync(); }
// only the lane-pair exchange (mask 1) is simulated
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int lane_mask) {
  ::cusim::PairCtx* c = ::cusim::pair_ctx();
  if (!c || lane_mask != 1) abort();
  const unsigned me = threadIdx.x & 1u;
  c->box[me] = v;
  c->sync();
  uint32_t r = c->box[me ^ 1u];
  c->sync();
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31));
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
  uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}

// ---- runtime API ----
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocPortable = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8,
       cudaFuncAttributePreferredSharedMemoryCarveout = 9, cudaSharedmemCarveoutMaxShared = 100 };
static inline const char* cudaGetErrorString(cudaError_t) { return "simulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = ::cusim::current_device(); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= ::cusim::device_count()) return 101; ::cusim::current_device() = d; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = ::cusim::device_count(); return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); if (*p) memset((void*)*p, 0xCD, n); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { *p = (T*)malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
template <class S> static inline cudaError_t cudaMemcpyToSymbol(S& sym, const void* s, size_t n) { memcpy((void*)&sym, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return cudaSuccess; }
