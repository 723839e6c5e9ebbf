// cuda_emu.h — the few CUDA built-ins the CTA-wide kernels use, emulated on the host: one OS thread per CUDA thread of a
// CTA, a pthread barrier for __syncthreads, per-warp barriers behind __ballot_sync / __shfl_xor_sync.  Include it BEFORE
// the device header under test.  Test infrastructure (tests/cpp/emu_*.cpp); see tests/test_ellcta_emulated.py.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace emu {
// ThreadSanitizer models pthread_barrier_wait as a release on entry and an acquire on exit of ONE sync object, so a thread
// that is already waiting in round k+1 would "publish" its accesses to a thread still leaving round k and hide a race
// between them.  Rotating over four barrier objects keeps the rounds apart (also across one deliberately skipped round).
constexpr int kRot = 4;
struct Cta {
    pthread_barrier_t all[kRot];
    std::vector<pthread_barrier_t> warp;   // kRot per warp
    std::vector<uint64_t> slot;            // 32 exchange slots per warp
};
thread_local Cta* cta = nullptr;
thread_local int tid = 0;
thread_local long warp_rounds = 0;
inline void warp_barrier() { pthread_barrier_wait(&cta->warp[(size_t)kRot * (tid >> 5) + (warp_rounds++ & (kRot - 1))]); }
inline uint64_t* slots() { return cta->slot.data() + 32 * (tid >> 5); }
}  // namespace emu

// EMU_DROP_BARRIER=<k> (self-test of the race detector): every thread skips its k-th __syncthreads, which keeps the
// barrier counts consistent but removes one ordering — ThreadSanitizer must then report a race
static const long emu_drop = std::getenv("EMU_DROP_BARRIER") ? std::atol(std::getenv("EMU_DROP_BARRIER")) : -1;
static thread_local long emu_barriers = 0;
static inline void __syncthreads() {
    const long k = emu_barriers++;
    if (k == emu_drop) return;
    pthread_barrier_wait(&emu::cta->all[k & (emu::kRot - 1)]);
#ifdef EMU_JITTER
    // a pseudo-random quarter of the threads leaves each barrier a little late: accesses that only a barrier orders are then
    // rarely simultaneous to the instruction, which is where ThreadSanitizer's shadow can miss them
    if ((((unsigned)emu::tid * 2654435761u + (unsigned)k * 40503u) >> 9) % 4 == 0) {
        volatile int spin = 0;
        for (int i = 0; i < 400; ++i) spin = spin + i;
    }
#endif
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    uint64_t* s = emu::slots();
    s[emu::tid & 31] = pred ? 1 : 0;
    emu::warp_barrier();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)s[l] << l;
    emu::warp_barrier();
    return m;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int off) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t* s = emu::slots();
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    s[emu::tid & 31] = bits;
    emu::warp_barrier();
    bits = s[(emu::tid & 31) ^ off];
    emu::warp_barrier();
    T out;
    std::memcpy(&out, &bits, sizeof(T));
    return out;
}
static inline double __ldg(const double* p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
using std::min;


namespace emu {
// runs fn(tid) on nt concurrent OS threads that form one emulated CTA
template <class F>
inline void run_cta(int nt, F fn) {
    Cta c;
    for (auto& b : c.all) pthread_barrier_init(&b, nullptr, nt);
    c.warp.resize((size_t)kRot * ((nt + 31) / 32));
    for (auto& w : c.warp) pthread_barrier_init(&w, nullptr, 32);
    c.slot.assign((size_t)32 * ((nt + 31) / 32), 0);
    std::vector<std::thread> ts;
    for (int t = 0; t < nt; ++t)
        ts.emplace_back([&c, &fn, t] {
            cta = &c;
            tid = t;
            warp_rounds = 0;
            fn(t);
        });
    for (auto& t : ts) t.join();
    for (auto& b : c.all) pthread_barrier_destroy(&b);
    for (auto& w : c.warp) pthread_barrier_destroy(&w);
}
}  // namespace emu
