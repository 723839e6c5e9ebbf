// cuda_emu.h — the few CUDA built-ins the CTA- and warp-wide kernels use, emulated on the host: one OS thread per CUDA thread
// of a CTA, a pthread barrier for __syncthreads, per-warp pthread barriers for __syncwarp, and a rendezvous the sanitizer
// cannot see behind __ballot_sync / __shfl*_sync.  Include it BEFORE the device code under test.
// Test infrastructure (tests/cpp/emu_*.cpp); see tests/test_emulated_kernels.py.
#pragma once
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

namespace emu {
// ThreadSanitizer models pthread_barrier_wait as a release on entry and an acquire on exit of ONE sync object, so a thread
// that is already waiting in round k+1 would "publish" its accesses to a thread still leaving round k and hide a race
// between them.  Rotating over four barrier objects keeps the rounds apart (also across one deliberately skipped round).
constexpr int kRot = 4;

// Shuffles and ballots exchange registers and order NO memory on the device.  Their emulation needs a rendezvous, but it must
// not hand ThreadSanitizer a happens-before edge (a missing __syncwarp next to a shuffle would hide behind it), so it is a
// sense-reversing barrier on plain atomics inside functions the sanitizer does not instrument.
#define EMU_QUIET __attribute__((no_sanitize("thread"), noinline))
struct QuietBarrier {
    unsigned count = 0, sense = 0;
};

struct Cta {
    pthread_barrier_t all[kRot];
    std::vector<pthread_barrier_t> warp;   // kRot per warp
    std::vector<QuietBarrier> quiet;       // one per warp
    std::vector<uint64_t> slot;            // 32 exchange slots per warp
};
thread_local Cta* cta = nullptr;
thread_local int tid = 0;
thread_local long warp_rounds = 0;
thread_local unsigned quiet_sense = 0;

inline void warp_barrier() { pthread_barrier_wait(&cta->warp[(size_t)kRot * (tid >> 5) + (warp_rounds++ & (kRot - 1))]); }

EMU_QUIET inline void quiet_wait() {
    QuietBarrier* b = &cta->quiet[tid >> 5];
    const unsigned my = quiet_sense ^= 1u;
    if (__atomic_add_fetch(&b->count, 1u, __ATOMIC_ACQ_REL) == 32u) {
        __atomic_store_n(&b->count, 0u, __ATOMIC_RELAXED);
        __atomic_store_n(&b->sense, my, __ATOMIC_RELEASE);
    } else {
        for (unsigned spins = 0; __atomic_load_n(&b->sense, __ATOMIC_ACQUIRE) != my; ++spins)
            if (spins > 32) sched_yield();
    }
}
// every lane deposits 8 bytes, then reads the deposit of lane `src` (or all 32, for a ballot)
EMU_QUIET inline uint64_t exchange(uint64_t mine, int src) {
    uint64_t* s = cta->slot.data() + 32 * (tid >> 5);
    s[tid & 31] = mine;
    quiet_wait();
    const uint64_t got = s[src & 31];
    quiet_wait();
    return got;
}
EMU_QUIET inline unsigned gather_bits(bool mine) {
    uint64_t* s = cta->slot.data() + 32 * (tid >> 5);
    s[tid & 31] = mine ? 1 : 0;
    quiet_wait();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)s[l] << l;
    quiet_wait();
    return m;
}
}  // namespace emu

// EMU_DROP_BARRIER=<k> / EMU_DROP_SYNCWARP=<k> (self-test of the race detector): every thread skips its k-th __syncthreads /
// __syncwarp, which keeps the barrier counts consistent but removes one ordering — ThreadSanitizer must then report a race
static const long emu_drop = std::getenv("EMU_DROP_BARRIER") ? std::atol(std::getenv("EMU_DROP_BARRIER")) : -1;
static const long emu_drop_warp = std::getenv("EMU_DROP_SYNCWARP") ? std::atol(std::getenv("EMU_DROP_SYNCWARP")) : -1;
static thread_local long emu_barriers = 0, emu_syncwarps = 0;

static inline void emu_jitter(long k) {
#ifdef EMU_JITTER
    // a pseudo-random quarter of the threads leaves each barrier a little late: accesses that only a barrier orders are then
    // rarely simultaneous to the instruction, which is where ThreadSanitizer's shadow can miss them
    if ((((unsigned)emu::tid * 2654435761u + (unsigned)k * 40503u) >> 9) % 4 == 0) {
        volatile int spin = 0;
        for (int i = 0; i < 400; ++i) spin = spin + i;
    }
#else
    (void)k;
#endif
}
static inline void emu_syncthreads_at(const char* file, int line) {
    const long k = emu_barriers++;
    if (k == emu_drop) {
        if (emu::tid == 0) std::fprintf(stderr, "[emu] dropped the __syncthreads at %s:%d\n", file, line);
        return;
    }
    pthread_barrier_wait(&emu::cta->all[k & (emu::kRot - 1)]);
    emu_jitter(k);
}
static inline void emu_syncwarp_at(const char* file, int line) {
    const long k = emu_syncwarps++;
    if (k == emu_drop_warp) {
        if (emu::tid == 0) std::fprintf(stderr, "[emu] dropped the __syncwarp at %s:%d\n", file, line);
        return;
    }
    emu::warp_barrier();
    emu_jitter(k + 7);
}
#define __syncthreads() emu_syncthreads_at(__FILE__, __LINE__)
#define __syncwarp(...) emu_syncwarp_at(__FILE__, __LINE__)

static inline unsigned __ballot_sync(unsigned, int pred) { return emu::gather_bits(pred != 0); }
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    bits = emu::exchange(bits, src);
    T out;
    std::memcpy(&out, &bits, sizeof(T));
    return out;
}
template <class T>
static inline T __shfl_xor_sync(unsigned m, T v, int off) { return __shfl_sync(m, v, (emu::tid & 31) ^ off); }

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;
static inline double __ldg(const double* p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {   // ((hi:lo) >> (shift & 31)), low word
    return (unsigned)(((((uint64_t)hi) << 32) | lo) >> (shift & 31));
}
static inline int __double2loint(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (int)(uint32_t)b; }
static inline int __double2hiint(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (int)(uint32_t)(b >> 32); }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double v;
    std::memcpy(&v, &b, 8);
    return v;
}
static inline double __ldcg(const double* p) { return *p; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_ACQ_REL); }
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_ACQ_REL); }
static inline long long clock64() { return (long long)std::clock(); }
static inline void __nanosleep(unsigned) { sched_yield(); }
#ifndef __grid_constant__
#define __grid_constant__
#endif
using std::min;

namespace emu {
// runs fn(tid) on nt concurrent OS threads that form one emulated CTA
template <class F>
inline void run_cta(int nt, F fn, EmuDim3 block = EmuDim3{0, 0, 0}, EmuDim3 grid = EmuDim3{1, 1, 1}) {
    Cta c;
    const size_t warps = (size_t)(nt + 31) / 32;
    for (auto& b : c.all) pthread_barrier_init(&b, nullptr, nt);
    c.warp.resize(kRot * warps);
    for (auto& w : c.warp) pthread_barrier_init(&w, nullptr, 32);
    c.quiet.resize(warps);
    c.slot.assign(32 * warps, 0);
    std::vector<std::thread> ts;
    for (int t = 0; t < nt; ++t)
        ts.emplace_back([&c, &fn, t, nt, block, grid] {
            cta = &c;
            tid = t;
            warp_rounds = 0;
            quiet_sense = 0;
            threadIdx.x = (unsigned)t; threadIdx.y = threadIdx.z = 0;
            blockIdx = block;
            gridDim = grid;
            blockDim.x = (unsigned)nt;
            fn(t);
        });
    for (auto& t : ts) t.join();
    for (auto& b : c.all) pthread_barrier_destroy(&b);
    for (auto& w : c.warp) pthread_barrier_destroy(&w);
}
}  // namespace emu
