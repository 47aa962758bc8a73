// mfsc_platform.cuh -- thin platform layer of the alignment library.
//
// The kernels in this directory are written warp-centric: a "warp" of MFSC_LANES lanes owns a
// unit of work (a probe tile, a read's match list, a DP band) and only warp-level primitives
// (shuffle / ballot / syncwarp) are used inside it.  Built by nvcc for sm_100a MFSC_LANES is 32
// and the primitives below are the hardware ones.  The same sources can be compiled by g++ with
// -DMFSC_EMU (tests/emu/, test infrastructure only): MFSC_LANES is then 1, every thread is its
// own warp and the launcher runs the grid sequentially.  That build exists so that the host
// orchestration and the per-warp control logic can be checked against the CPU oracle on a
// machine without a GPU; the Python package never loads it (see metafinishersc_b200/_lib.py).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifdef MFSC_EMU
// ------------------------------------------------------------------------------------------
// host emulation (tests only)
// ------------------------------------------------------------------------------------------
#include <algorithm>
#include <vector>
#define MFSC_LANES 1
#define MFSC_DEV
#define MFSC_KERNEL static void
#define MFSC_KERNEL_LB(T, B) static void
#define MFSC_HD
#define MFSC_FORCEINLINE inline
#define MFSC_NOINLINE static
struct emu_dim3 { unsigned x, y, z; };
extern thread_local emu_dim3 emu_threadIdx, emu_blockIdx, emu_blockDim, emu_gridDim;
#define threadIdx emu_threadIdx
#define blockIdx emu_blockIdx
#define blockDim emu_blockDim
#define gridDim emu_gridDim
typedef int cudaStream_t;
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
static inline int2 make_int2(int x, int y) { int2 v; v.x = x; v.y = y; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

#define MFSC_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
  do {                                                                                        \
    emu_gridDim = emu_dim3{(unsigned)(grid), 1, 1};                                           \
    emu_blockDim = emu_dim3{(unsigned)(block), 1, 1};                                         \
    for (unsigned _b = 0; _b < (unsigned)(grid); _b++)                                        \
      for (unsigned _t = 0; _t < (unsigned)(block); _t++) {                                   \
        emu_blockIdx = emu_dim3{_b, 0, 0};                                                    \
        emu_threadIdx = emu_dim3{_t, 0, 0};                                                   \
        kernel(__VA_ARGS__);                                                                  \
      }                                                                                       \
  } while (0)

static inline void wsync() {}
static inline int wmax_i32(int v) { return v; }
static inline int wmin_i32(int v) { return v; }
static inline long long wmax_i64(long long v) { return v; }
static inline long long wmin_i64(long long v) { return v; }
static inline int wsum_i32(int v) { return v; }
static inline long long wsum_i64(long long v) { return v; }
static inline unsigned wballot(bool p) { return p ? 1u : 0u; }
static inline unsigned wmatch_any_i32(int) { return 1u; }
static inline int wbcast_i32(int v, int) { return v; }
static inline long long wbcast_i64(long long v, int) { return v; }
static inline unsigned atomic_add_u32(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomic_add_u64(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p; *p = o + v; return o;
}
static inline int atomic_add_i32(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline int atomic_max_i32(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomic_cas_i32(int* p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned wbcast_u32(unsigned v, int) { return v; }
static inline int wshfl_i32(int v, int) { return v; }
static inline int dev_ffs32(unsigned x) { return x ? __builtin_ctz(x) + 1 : 0; }
static inline int dev_ffs64(unsigned long long x) { return x ? __builtin_ctzll(x) + 1 : 0; }
static inline int dev_clz32(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int dev_clz64(unsigned long long x) { return x ? __builtin_clzll(x) : 64; }
static inline int dev_popc32(unsigned x) { return __builtin_popcount(x); }
static inline unsigned dev_byte_rev(unsigned x) { return __builtin_bswap32(x); }
static inline unsigned dev_funnel_r(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (sh & 31u)); }
static inline unsigned long long dev_brev64(unsigned long long x) {
  unsigned long long r = 0;
  for (int i = 0; i < 64; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}
static inline unsigned ld_volatile_u32(const volatile unsigned* p) { return *p; }
static inline int ld_volatile_i32(const volatile int* p) { return *p; }
static inline unsigned long long ld_volatile_u64(const volatile unsigned long long* p) { return *p; }
static inline void st_volatile_u64(volatile unsigned long long* p, unsigned long long v) { *p = v; }
static inline void dev_threadfence() {}
template <class T> static inline T ldg(const T* p) { return *p; }
template <class T> static inline T ld_stream(const T* p) { return *p; }
template <class T> static inline void st_stream(T* p, T v) { *p = v; }

#else
// ------------------------------------------------------------------------------------------
// sm_100a
// ------------------------------------------------------------------------------------------
#include <cuda_runtime.h>
#define MFSC_LANES 32
#define MFSC_DEV __device__ __forceinline__
#define MFSC_KERNEL __global__ void
#define MFSC_KERNEL_LB(T, B) __global__ void __launch_bounds__(T, B)
#define MFSC_HD __host__ __device__
#define MFSC_FORCEINLINE __forceinline__
#define MFSC_NOINLINE __device__ __noinline__

#define MFSC_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
  kernel<<<(unsigned)(grid), (unsigned)(block), (size_t)(smem), stream>>>(__VA_ARGS__)

#define MFSC_FULL 0xffffffffu
MFSC_DEV void wsync() { __syncwarp(); }
MFSC_DEV int wmax_i32(int v) { return __reduce_max_sync(MFSC_FULL, v); }
MFSC_DEV int wmin_i32(int v) { return __reduce_min_sync(MFSC_FULL, v); }
MFSC_DEV int wsum_i32(int v) { return __reduce_add_sync(MFSC_FULL, v); }
MFSC_DEV long long wmax_i64(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_xor_sync(MFSC_FULL, v, o); v = t > v ? t : v; }
  return v;
}
MFSC_DEV long long wmin_i64(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_xor_sync(MFSC_FULL, v, o); v = t < v ? t : v; }
  return v;
}
MFSC_DEV long long wsum_i64(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MFSC_FULL, v, o);
  return v;
}
MFSC_DEV unsigned wballot(bool p) { return __ballot_sync(MFSC_FULL, p); }
MFSC_DEV unsigned wmatch_any_i32(int v) { return __match_any_sync(MFSC_FULL, v); }   // lanes holding the same value
MFSC_DEV int wbcast_i32(int v, int src) { return __shfl_sync(MFSC_FULL, v, src); }
MFSC_DEV long long wbcast_i64(long long v, int src) { return __shfl_sync(MFSC_FULL, v, src); }
MFSC_DEV unsigned atomic_add_u32(unsigned* p, unsigned v) { return atomicAdd(p, v); }
MFSC_DEV unsigned long long atomic_add_u64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
MFSC_DEV int atomic_add_i32(int* p, int v) { return atomicAdd(p, v); }
MFSC_DEV int atomic_max_i32(int* p, int v) { return atomicMax(p, v); }
MFSC_DEV int atomic_cas_i32(int* p, int cmp, int v) { return atomicCAS(p, cmp, v); }
MFSC_DEV unsigned wbcast_u32(unsigned v, int src) { return __shfl_sync(MFSC_FULL, v, src); }
MFSC_DEV int wshfl_i32(int v, int src) { return __shfl_sync(MFSC_FULL, v, src); }      // per-lane source
MFSC_DEV int dev_ffs32(unsigned x) { return __ffs((int)x); }
MFSC_DEV int dev_ffs64(unsigned long long x) { return __ffsll((long long)x); }
MFSC_DEV int dev_clz32(unsigned x) { return __clz((int)x); }
MFSC_DEV int dev_clz64(unsigned long long x) { return __clzll((long long)x); }
MFSC_DEV int dev_popc32(unsigned x) { return __popc(x); }
MFSC_DEV unsigned dev_byte_rev(unsigned x) { return __byte_perm(x, 0u, 0x0123u); }
MFSC_DEV unsigned dev_funnel_r(unsigned lo, unsigned hi, unsigned sh) { return __funnelshift_r(lo, hi, sh); }
MFSC_DEV unsigned long long dev_brev64(unsigned long long x) { return __brevll(x); }
MFSC_DEV unsigned ld_volatile_u32(const volatile unsigned* p) { return *p; }
MFSC_DEV int ld_volatile_i32(const volatile int* p) { return *p; }
MFSC_DEV unsigned long long ld_volatile_u64(const volatile unsigned long long* p) { return *p; }
MFSC_DEV void st_volatile_u64(volatile unsigned long long* p, unsigned long long v) { *p = v; }
MFSC_DEV void dev_threadfence() { __threadfence(); }
template <class T> MFSC_DEV T ldg(const T* p) { return __ldg(p); }
// streaming load: evict-first in L2, so a one-touch random stream does not displace the reused tables
template <class T> MFSC_DEV T ld_stream(const T* p) { return __ldcs(p); }
template <class T> MFSC_DEV void st_stream(T* p, T v) { __stcs(p, v); }
#endif

// common helpers --------------------------------------------------------------------------
MFSC_DEV int lane_id() { return (int)(threadIdx.x % MFSC_LANES); }
MFSC_DEV int warp_in_block() { return (int)(threadIdx.x / MFSC_LANES); }
MFSC_DEV int warps_per_block() { return (int)(blockDim.x / MFSC_LANES); }
MFSC_DEV long long warp_global() { return (long long)blockIdx.x * warps_per_block() + warp_in_block(); }
MFSC_DEV long long warps_total() { return (long long)gridDim.x * warps_per_block(); }
MFSC_DEV long long thread_global() { return (long long)blockIdx.x * blockDim.x + threadIdx.x; }
MFSC_DEV long long threads_total() { return (long long)gridDim.x * blockDim.x; }
