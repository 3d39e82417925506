// cuda_emu.h -- TEST INFRASTRUCTURE: a tiny host emulation of the CUDA constructs kf_kernels.cuh uses,
// so the kernel LOGIC (decode, line-state resolution, byte walker, tiling, fold) can be exercised on a
// box without a GPU.  One OS thread per CUDA thread, pthread barriers for warp collectives and
// __syncthreads.  It is slow and only meant for kilobyte-to-megabyte inputs in tests/.
// It is never part of the product: libkfcount.so is built by nvcc from the same header without KF_EMU.
#pragma once
#include <pthread.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define KF_NOINLINE

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct ulonglong2 { unsigned long long x, y; };
inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { return ulonglong2{a, b}; }
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; };

namespace emu {
struct WarpShared {
    pthread_barrier_t bar;
    uint64_t slot[32];
};
struct BlockShared {
    pthread_barrier_t bar;
    std::vector<uint8_t> smem;
    std::vector<WarpShared> warps;
};
extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern thread_local BlockShared *t_block;
inline WarpShared &warp() { return t_block->warps[t_threadIdx.x >> 5]; }
inline uint64_t exchange(uint64_t v, int src) {
    WarpShared &w = warp();
    w.slot[t_threadIdx.x & 31] = v;
    pthread_barrier_wait(&w.bar);
    uint64_t r = w.slot[src & 31];
    pthread_barrier_wait(&w.bar);
    return r;
}
// run kernel body `fn` for grid x block threads (block must be a multiple of 32)
void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &fn);
}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::t_blockDim
#define gridDim emu::t_gridDim
#define KF_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::t_block->smem.data())

inline void __syncthreads() { pthread_barrier_wait(&emu::t_block->bar); }
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::WarpShared &w = emu::warp();
    w.slot[threadIdx.x & 31] = pred ? 1 : 0;
    pthread_barrier_wait(&w.bar);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)(w.slot[i] & 1) << i;
    pthread_barrier_wait(&w.bar);
    return r;
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return (T)emu::exchange((uint64_t)v, src); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m) { return (T)emu::exchange((uint64_t)v, (int)((threadIdx.x & 31) ^ (unsigned)m)); }
template <typename T>
inline T __ldg(const T *p) { return *p; }
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    uint64_t pool = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t b = (uint32_t)(pool >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;
        r |= b << (8 * i);
    }
    return r;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline void __threadfence() {}
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline uint32_t __brev(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned long long __brevll(unsigned long long x) {
    return ((unsigned long long)__brev((uint32_t)x) << 32) | __brev((uint32_t)(x >> 32));
}
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31;
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicOr(uint32_t *p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicSub(uint32_t *p, uint32_t v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }
#define __shared__ static
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int d) {
    int l = (int)(threadIdx.x & 31);
    T r = (T)emu::exchange((uint64_t)v, l - d < 0 ? l : l - d);
    return r;
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int d) {
    int l = (int)(threadIdx.x & 31);
    T r = (T)emu::exchange((uint64_t)v, l + d > 31 ? l : l + d);
    return r;
}
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
