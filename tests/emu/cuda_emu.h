// Functional CPU emulation of the small CUDA subset the radb kernels use.
// TEST INFRASTRUCTURE ONLY: it lets `tests/` run the *same kernel source*
// (multimodal-isic_b200/csrc/radb_kernels.cuh) on a GPU-less box to check kernel
// logic against the oracle.  The product never links or loads this.
//
// Model: one OS thread per CUDA thread of a CTA; __syncthreads / warp collectives are
// real barriers, atomics are real atomics.  Warp collectives must be called by all 32
// lanes (full mask, convergent) -- the kernels are written that way.
#pragma once
#define RADB_EMU 1
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)

struct alignas(16) uint4 {
    unsigned x, y, z, w;
};
struct alignas(16) int4 {
    int x, y, z, w;
};

namespace emu {
struct Warp {
    std::barrier<> bar{32};
    std::barrier<> gbar[4] = {std::barrier<>(8), std::barrier<>(8), std::barrier<>(8), std::barrier<>(8)};
    uint64_t slot[32];
    // full mask -> the warp barrier; 0xff << 8g -> the barrier of the aligned 8-lane group g (sub-warp collectives)
    std::barrier<>& pick(unsigned mask) {
        if (mask == 0xffffffffu) return bar;
        for (int g = 0; g < 4; g++)
            if (mask == (0xffu << (8 * g))) return gbar[g];
        std::fprintf(stderr, "cuda_emu: unsupported collective mask %08x\n", mask);
        std::abort();
    }
};
struct Ctx {
    unsigned tid = 0, bid = 0, nthreads = 0;
    std::barrier<>* cta = nullptr;
    Warp* warp = nullptr;
};
inline thread_local Ctx ctx;
struct Dim3 {
    unsigned x, y, z;
};
inline Dim3 tidx() { return {ctx.tid, 0, 0}; }
inline Dim3 bidx() { return {ctx.bid, 0, 0}; }
inline Dim3 bdim() { return {ctx.nthreads, 1, 1}; }

template <class T>
inline uint64_t bits(T v) {
    uint64_t b = 0;
    static_assert(sizeof(T) <= 8, "");
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T unbits(uint64_t b) {
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}
// every lane publishes v, then reads the value of lane `src`
template <class T>
inline T exchange(T v, int src, unsigned mask = 0xffffffffu) {
    Warp* w = ctx.warp;
    int lane = ctx.tid & 31;
    std::barrier<>& b = w->pick(mask);
    w->slot[lane] = bits(v);
    b.arrive_and_wait();
    T r = unbits<T>(w->slot[src & 31]);
    b.arrive_and_wait();
    return r;
}
// every lane publishes v and gets all 32 values
template <class T>
inline void gather(T v, T out[32], unsigned mask = 0xffffffffu) {
    Warp* w = ctx.warp;
    int lane = ctx.tid & 31;
    std::barrier<>& b = w->pick(mask);
    w->slot[lane] = bits(v);
    b.arrive_and_wait();
    for (int i = 0; i < 32; i++) out[i] = unbits<T>(w->slot[i]);  // only the lanes in `mask` are meaningful
    b.arrive_and_wait();
}

// launch: grid CTAs run one after another, `nthreads` OS threads each
template <class F>
void launch(unsigned grid, unsigned nthreads, F body) {
    std::barrier<> cta(nthreads);
    std::vector<std::unique_ptr<Warp>> warps;
    for (unsigned i = 0; i < (nthreads + 31) / 32; i++) warps.emplace_back(new Warp());
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; t++)
        th.emplace_back([&, t]() {
            ctx.tid = t;
            ctx.nthreads = nthreads;
            ctx.cta = &cta;
            ctx.warp = warps[t / 32].get();
            for (unsigned b = 0; b < grid; b++) {
                ctx.bid = b;
                body();
                cta.arrive_and_wait();
            }
        });
    for (auto& t : th) t.join();
}
}  // namespace emu

#define threadIdx (emu::tidx())
#define blockIdx (emu::bidx())
#define blockDim (emu::bdim())

inline void __syncthreads() { emu::ctx.cta->arrive_and_wait(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::ctx.warp->pick(mask).arrive_and_wait(); }

template <class T>
inline T __shfl_sync(unsigned m, T v, int src) { return emu::exchange(v, src, m); }
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int m) { return emu::exchange(v, (emu::ctx.tid & 31) ^ m, mask); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    int lane = emu::ctx.tid & 31;
    T r = emu::exchange(v, lane - (int)d < 0 ? lane : lane - (int)d);
    return r;
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    int lane = emu::ctx.tid & 31;
    return emu::exchange(v, lane + (int)d > 31 ? lane : lane + (int)d);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    int all[32];
    emu::gather<int>(pred ? 1 : 0, all, mask);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) r |= (all[i] ? 1u : 0u) << i;
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline unsigned __match_any_sync(unsigned, unsigned v) {
    unsigned all[32];
    emu::gather<unsigned>(v, all);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (all[i] == v ? 1u : 0u) << i;
    return r;
}
inline double __dmul_rn(double a, double b) {
    volatile double r = a * b;  // a separately rounded product even when the harness is built with FMA contraction
    return r;
}
inline float __fadd_rn(float a, float b) {
    volatile float r = a + b;
    return r;
}
inline double __dadd_rn(double a, double b) {
    volatile double r = a + b;
    return r;
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }

template <class T>
inline T atomicAdd(T* a, T v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
inline int atomicMin(int* a, int v) {
    int old = __atomic_load_n(a, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int atomicMax(int* a, int v) {
    int old = __atomic_load_n(a, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
template <class T>
inline T atomicCAS(T* a, T cmp, T val) {
    __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
