// stream_ptx_emu.h — takes the place of ellalgo_rs_b200/csrc/stream_ptx.cuh in the host emulation: the same functions
// (mbarrier, TMA tensor / bulk copies, flagged 16-byte lines, programmatic dependent launch) without PTX.  Include after
// cuda_emu.h and before the kernel source.
//   * A TMA copy runs at once in the issuing thread and then completes its bytes on the mbarrier; the mbarrier is a 64-bit
//     word {phase, pending arrivals, outstanding bytes} written with release and polled with acquire — which is the ordering
//     the hardware gives (data of a completed phase is visible to whoever observed the phase), so ThreadSanitizer checks
//     exactly the hand-off the kernel relies on.  That the copies are asynchronous on the device (a tile being stored while its
//     stage is refilled) is NOT modelled.
//   * A flagged line is two 8-byte atomics {lo32 | flag << 32}, {hi32 | flag << 32}: 8-byte stores are never torn.
//   * CUtensorMap: the emulation keeps {base, rows, cols, ld, box} in the opaque 128 bytes (emu_make_map).
#pragma once
#define ELLCUDA_STREAM_PTX_CUH   // keeps the device header out
#include <cuda.h>

#include "../../ellalgo_rs_b200/csrc/common.cuh"

namespace ellcuda {
namespace {

struct EmuMap {
    double* base;
    long long rows, cols, ld;
    int box_rows, box_cols;
};
static_assert(sizeof(EmuMap) <= sizeof(CUtensorMap), "");
inline void emu_make_map(CUtensorMap* map, double* base, long long rows, long long cols, long long ld, int box_rows, int box_cols) {
    std::memset(map, 0, sizeof *map);
    const EmuMap m{base, rows, cols, ld, box_rows, box_cols};
    std::memcpy(map, &m, sizeof m);
}
inline EmuMap emu_map(const CUtensorMap* map) {
    EmuMap m;
    std::memcpy(&m, map, sizeof m);
    return m;
}

// ---- mbarrier: bit 63 phase | bits 32..47 pending arrivals | bits 48..62 arrival count | bits 0..31 outstanding bytes
inline unsigned long long mb_pack(unsigned phase, unsigned count, unsigned pending, unsigned tx) {
    return ((unsigned long long)phase << 63) | ((unsigned long long)count << 48) | ((unsigned long long)pending << 32) | tx;
}
inline void mb_update(void* bar, int arrive, long long tx_delta) {   // one writer at a time (the CTA's elected thread)
    unsigned long long* w = static_cast<unsigned long long*>(bar);
    const unsigned long long v = __atomic_load_n(w, __ATOMIC_RELAXED);
    unsigned phase = (unsigned)(v >> 63), count = (unsigned)((v >> 48) & 0x7fff), pending = (unsigned)((v >> 32) & 0xffff);
    long long tx = (long long)(unsigned)(v & 0xffffffffu) + tx_delta;
    pending -= (unsigned)arrive;
    if (pending == 0 && tx == 0) { phase ^= 1u; pending = count; }
    __atomic_store_n(w, mb_pack(phase, count, pending, (unsigned)tx), __ATOMIC_RELEASE);
}
inline void mbar_init(void* bar, unsigned count) { __atomic_store_n(static_cast<unsigned long long*>(bar), mb_pack(0, count, count, 0), __ATOMIC_RELEASE); }
inline void mbar_fence_init() {}
inline void mbar_expect_tx(void* bar, unsigned bytes) { mb_update(bar, 1, (long long)bytes); }
inline void mbar_wait(void* bar, unsigned parity) {
    const unsigned long long* w = static_cast<const unsigned long long*>(bar);
    for (unsigned spins = 0; (unsigned)(__atomic_load_n(w, __ATOMIC_ACQUIRE) >> 63) == parity; ++spins)
        if (spins > 32) sched_yield();
}
inline void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, void* bar) {   // c0 = column, c1 = row; zero fill outside
    const EmuMap m = emu_map(map);
    double* d = static_cast<double*>(dst);
    for (int r = 0; r < m.box_rows; ++r)
        for (int c = 0; c < m.box_cols; ++c) {
            const long long row = (long long)c1 + r, col = (long long)c0 + c;
            d[r * m.box_cols + c] = (row < m.rows && col < m.cols) ? m.base[row * m.ld + col] : 0.0;
        }
    mb_update(bar, 0, -(long long)m.box_rows * m.box_cols * 8);
}
inline void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {          // clipped at the tensor's edges
    const EmuMap m = emu_map(map);
    const double* s = static_cast<const double*>(src);
    for (int r = 0; r < m.box_rows; ++r)
        for (int c = 0; c < m.box_cols; ++c) {
            const long long row = (long long)c1 + r, col = (long long)c0 + c;
            if (row < m.rows && col < m.cols) m.base[row * m.ld + col] = s[r * m.box_cols + c];
        }
}
inline void bulk_load_1d(void* dst, const void* src, unsigned bytes, void* bar) {
    std::memcpy(dst, src, bytes);
    mb_update(bar, 0, -(long long)bytes);
}
inline void bulk_store_1d(void* dst, const void* src, unsigned bytes) { std::memcpy(dst, src, bytes); }
inline void bulk_commit() {}
template <int N>
inline void bulk_wait_read() {}
inline void bulk_wait_all() {}
inline void fence_async_smem() {}

// ---- flagged lines
inline void ll_store(uint4* line, double v, unsigned flag) {
    unsigned long long* h = reinterpret_cast<unsigned long long*>(line);
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    __atomic_store_n(h, ((unsigned long long)flag << 32) | lo, __ATOMIC_RELEASE);
    __atomic_store_n(h + 1, ((unsigned long long)flag << 32) | hi, __ATOMIC_RELEASE);
}
inline void ll_peek(const uint4* line, unsigned (&w)[4]) {
    const unsigned long long* h = reinterpret_cast<const unsigned long long*>(line);
    const unsigned long long a = __atomic_load_n(h, __ATOMIC_ACQUIRE), b = __atomic_load_n(h + 1, __ATOMIC_ACQUIRE);
    w[0] = (unsigned)a; w[1] = (unsigned)(a >> 32); w[2] = (unsigned)b; w[3] = (unsigned)(b >> 32);
}
inline bool ll_load(const uint4* line, unsigned flag, double& v) {
    unsigned w[4];
    for (long spins = 0;; ++spins) {
        ll_peek(line, w);
        if (w[1] == flag && w[3] == flag) break;
        if (spins > 200000000L) return false;
        if (spins > 32) sched_yield();
    }
    v = __hiloint2double((int)w[2], (int)w[0]);
    return true;
}
inline void pdl_wait() {}
inline void pdl_trigger() {}

}  // namespace
}  // namespace ellcuda
