// cuda_runtime.h (shim) — TEST INFRASTRUCTURE.  Found instead of the CUDA header when the kernel
// sources are compiled by g++ for the warp emulator (tests/emu/warp_emu.hpp): CUDA's function
// qualifiers vanish, the device intrinsics the kernels use become plain C++ (same IEEE results:
// no contraction, correctly rounded sqrt and divide), and the warp primitives become rendez-vous of
// the 32 lane fibers, tagged with their source line.
#pragma once
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <cmath>

#include <functional>

#include "../warp_emu.hpp"

// generic kernel launch of the emulated build (TESS_LAUNCH in common.cuh); defined by the harness
void emu_launch_generic(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body);

#define TESS_WARP_EMU 1
#define TESS_UNIFORM_BEGIN(ptr, bytes) emu::uniform_begin((ptr), (bytes), __LINE__)
#define TESS_UNIFORM_END() emu::uniform_end(__LINE__)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static const
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__ __restrict

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
static const int cudaSuccess = 0;
inline const char* cudaGetErrorString(cudaError_t) { return "cuda call in the warp emulator"; }

struct __attribute__((aligned(16))) double2 {
    double x, y;
};
struct __attribute__((aligned(16))) double4 {
    double x, y, z, w;
};
struct __attribute__((aligned(8))) uint2 {
    unsigned x, y;
};
struct __attribute__((aligned(16))) uint4 {
    unsigned x, y, z, w;
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }

#define threadIdx (emu::thread_idx())
#define blockIdx (emu::block_idx())
#define blockDim (emu::block_dim())
#define gridDim (emu::grid_dim())

// ---- warp primitives ------------------------------------------------------------------------
#define __shfl_sync(m, v, s) emu::shfl((m), (v), (int)(s), __LINE__)
#define __shfl_up_sync(m, v, d) emu::shfl_up((m), (v), (unsigned)(d), __LINE__)
#define __shfl_down_sync(m, v, d) emu::shfl_down((m), (v), (unsigned)(d), __LINE__)
#define __shfl_xor_sync(m, v, d) emu::shfl_xor((m), (v), (unsigned)(d), __LINE__)
#define __ballot_sync(m, p) emu::ballot((m), (bool)(p), __LINE__)
#define __all_sync(m, p) emu::all_((m), (bool)(p), __LINE__)
#define __any_sync(m, p) emu::any_((m), (bool)(p), __LINE__)
#define __reduce_max_sync(m, v) emu::reduce_max((m), (unsigned)(v), __LINE__)
#define __reduce_min_sync(m, v) emu::reduce_min((m), (unsigned)(v), __LINE__)
#define __reduce_or_sync(m, v) emu::reduce_or((m), (unsigned)(v), __LINE__)
#define __reduce_add_sync(m, v) emu::reduce_add((m), (unsigned)(v), __LINE__)
#define __match_any_sync(m, v) emu::match_any((m), (uint64_t)(v), __LINE__)
#define __syncwarp() emu::syncwarp(__LINE__)
#define __syncthreads() emu::syncthreads(__LINE__)
// a spin-wait gives the other warps of the block their turn
#define __nanosleep(ns) emu::yield_to_scheduler()
#define __threadfence_block() ((void)0)
// static shared arrays: one block at a time runs on a host thread
#define __shared__ static thread_local

// ---- scalar intrinsics ----------------------------------------------------------------------
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline long long __double_as_longlong(double v) {
    long long r;
    memcpy(&r, &v, 8);
    return r;
}
inline double __longlong_as_double(long long v) {
    double r;
    memcpy(&r, &v, 8);
    return r;
}
inline int __double2hiint(double v) { return (int)(__double_as_longlong(v) >> 32); }
inline int __double2loint(double v) { return (int)(__double_as_longlong(v) & 0xffffffffll); }
inline double __hiloint2double(int hi, int lo) { return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo)); }
inline unsigned __double2uint_rz(double v) {  // saturating, NaN -> 0
    if (!(v > 0.0)) return 0u;
    if (v >= 4294967295.0) return 0xffffffffu;
    return (unsigned)v;
}
// per-byte unsigned a < b -> 0xff
inline unsigned __vcmpltu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; ++i)
        if (((a >> (8 * i)) & 0xffu) < ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
}
inline unsigned __vcmpgtu4(unsigned a, unsigned b) { return __vcmpltu4(b, a); }
inline unsigned __vcmpeq4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; ++i)
        if (((a >> (8 * i)) & 0xffu) == ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)((v >> (8 * (sel & 7u))) & 0xffu);
        if (sel & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31u));
}
template <class T>
inline T __ldg(const T* p) { return *p; }

template <class T>
inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T>
inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T>
inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <class T>
inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T>
inline T min(T a, T b) { return b < a ? b : a; }
template <class T>
inline T max(T a, T b) { return a < b ? b : a; }
// host-side runtime calls the launchers make
inline cudaError_t cudaMallocAsync(void* pp, size_t bytes, cudaStream_t) { *static_cast<void**>(pp) = malloc(bytes); return cudaSuccess; }
template <class T>
inline cudaError_t cudaMallocAsync(T** pp, size_t bytes, cudaStream_t) { *pp = static_cast<T*>(malloc(bytes)); return cudaSuccess; }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <class T>
inline T atomicMax(T* p, T v) {
    T o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
