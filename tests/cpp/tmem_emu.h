// tmem_emu.h — tensor memory for the host emulation: takes the place of ellalgo_rs_b200/csrc/tmem.cuh (same functions, same
// addressing: taddr = lane << 16 | column, a warp reaches the 32 lanes starting at its lane field, N consecutive 32-bit
// columns per thread) on a plain 128 x 512 array of 32-bit cells.  Include after cuda_emu.h and before the kernel source.
// tcgen05.ld / .st are asynchronous on the device (tm_wait_*); here they complete at once, so a missing wait is not caught —
// what IS caught, under ThreadSanitizer, is two warps touching the same cell without a barrier in between.
#pragma once
#define ELLCUDA_TMEM_CUH   // keeps the device header out
#include <cstdint>

namespace ellcuda {
static uint32_t emu_tmem[128][512];

template <int N>
static inline void tm_ld(uint32_t taddr, uint32_t (&r)[N]) {
    const uint32_t lane = (taddr >> 16) + (uint32_t)(emu::tid & 31), col = taddr & 0xffffu;
    for (int i = 0; i < N; ++i) r[i] = emu_tmem[lane][col + i];
}
template <int N>
static inline void tm_st(uint32_t taddr, const uint32_t (&r)[N]) {
    const uint32_t lane = (taddr >> 16) + (uint32_t)(emu::tid & 31), col = taddr & 0xffffu;
    for (int i = 0; i < N; ++i) emu_tmem[lane][col + i] = r[i];
}
static inline void tm_ld16(uint32_t taddr, uint32_t (&r)[16]) { tm_ld<16>(taddr, r); }
static inline void tm_ld32(uint32_t taddr, uint32_t (&r)[32]) { tm_ld<32>(taddr, r); }
static inline void tm_st16(uint32_t taddr, const uint32_t (&r)[16]) { tm_st<16>(taddr, r); }
static inline void tm_st32(uint32_t taddr, const uint32_t (&r)[32]) { tm_st<32>(taddr, r); }
static inline void tm_wait_ld() {}
static inline void tm_wait_st() {}
#define tm_fence_sync() __syncthreads()
static inline double mk_double(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }
#define tm_alloc_all(slot, warp) (tm_fence_sync(), (void)(slot), (void)(warp), 0u)
#define tm_free_all(base, warp) ((void)(base), (void)(warp), tm_fence_sync())
}  // namespace ellcuda
