// Micro-benchmark: can tensor memory (TMEM) serve as a second read-only table store next to shared memory for FP64 data?
// The lowpass oracle kernel is bound by the shared-memory data pipe (profiles/r01c_*): this measures, per SM, the bytes per
// clock of (a) conflict-free 64-bit LDS, (b) tcgen05.ld 32x32b.x32 (16 doubles per lane), (c) both interleaved, each
// feeding one DFMA per double as the oracle's row-dot does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_tmem tools/ubench_tmem.cu && ./ubench_tmem
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kWarps = 16;
constexpr int kIters = 4096;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
        :: "r"(taddr),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ double mk(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

// MODE 0: shared only (32 doubles per round), 1: TMEM only (32 doubles per round, two x32 loads),
// 2: 16 doubles from TMEM + 16 from shared per round.
template <int MODE>
__global__ void __launch_bounds__(kWarps * 32, 1) bench(double* out, long long* cycles, int check) {
    extern __shared__ __align__(16) double tab[];   // 480 rows x 33
    __shared__ uint32_t tmem_base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 480 * 33; i += blockDim.x) tab[i] = 1.0 + 1e-3 * (i % 977);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n"
                     :: "l"((uint64_t)__cvta_generic_to_shared(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tbase = tmem_base_s;
    const uint32_t tlane = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
    // fill: the first warp of each lane quarter writes 512 columns: value(lane, col pair c) = 1 + 1e-3*((lane*256+c) % 977)
    if (warp < 4) {
        for (int blk = 0; blk < 16; ++blk) {
            uint32_t r[32];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const double v = 1.0 + 1e-3 * (((32 * warp + lane) * 256 + blk * 16 + q) % 977);
                r[2 * q] = (uint32_t)__double2loint(v);
                r[2 * q + 1] = (uint32_t)__double2hiint(v);
            }
            tmem_st32(tlane + blk * 32, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");

    if (check) {  // read back block 3 and compare
        uint32_t r[32];
        tmem_ld32(tlane + 3 * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        double bad = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const double v = 1.0 + 1e-3 * (((32 * (warp & 3) + lane) * 256 + 3 * 16 + q) % 977);
            if (mk(r[2 * q], r[2 * q + 1]) != v) bad += 1.0;
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = bad;
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tbase));
        return;
    }

    double x[4] = {1.0000001, 0.9999999, 1.0000002, 0.9999998};
    double acc0 = 0.0, acc1 = 0.0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
        const int k = (it * 7 + warp * 29) % 448;          // a moving 32-row window, like the oracle's cursor
        const double* row = tab + (k + lane) * 33;
        const uint32_t blk = (uint32_t)((it + warp) & 7);  // which 16-double block pair of this lane's TMEM columns
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                acc0 += row[j] * x[j & 3];
                acc1 += row[j + 1] * x[(j + 1) & 3];
            }
        } else if (MODE == 1) {
            uint32_t r[32], s[32];
            tmem_ld32(tlane + blk * 64, r);
            tmem_ld32(tlane + blk * 64 + 32, s);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
                acc0 += mk(r[2 * q], r[2 * q + 1]) * x[q & 3];
                acc1 += mk(r[2 * q + 2], r[2 * q + 3]) * x[(q + 1) & 3];
            }
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
                acc0 += mk(s[2 * q], s[2 * q + 1]) * x[q & 3];
                acc1 += mk(s[2 * q + 2], s[2 * q + 3]) * x[(q + 1) & 3];
            }
        } else {
            uint32_t r[32];
            tmem_ld32(tlane + blk * 64, r);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                acc0 += row[j] * x[j & 3];
                acc1 += row[j + 1] * x[(j + 1) & 3];
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
                acc0 += mk(r[2 * q], r[2 * q + 1]) * x[q & 3];
                acc1 += mk(r[2 * q + 2], r[2 * q + 3]) * x[(q + 1) & 3];
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tbase));
}

template <int MODE>
void run(const char* name, double* out, long long* cyc, int sms) {
    const size_t smem = 480 * 33 * sizeof(double);
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<sms, kWarps * 32, smem>>>(out, cyc, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<MODE><<<sms, kWarps * 32, smem>>>(out, cyc, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c0 = 0;
    cudaMemcpy(&c0, cyc, sizeof(c0), cudaMemcpyDeviceToHost);
    const double bytes = (double)kIters * kWarps * 32 * 32 * 8;   // per SM
    printf("%-28s %s  %.3f ms  cycles/CTA %lld  -> %.1f B/clk/SM (doubles consumed), %.2f DFMA/clk/SM\n", name,
           cudaGetErrorString(err), ms, c0, bytes / (double)c0, bytes / 8 / (double)c0);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * kWarps * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const size_t smem = 480 * 33 * sizeof(double);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<1><<<sms, kWarps * 32, smem>>>(out, cyc, 1);
    cudaError_t err = cudaDeviceSynchronize();
    double* h = new double[sms * kWarps * 32];
    cudaMemcpy(h, out, sizeof(double) * sms * kWarps * 32, cudaMemcpyDeviceToHost);
    double bad = 0;
    for (int i = 0; i < sms * kWarps * 32; ++i) bad += h[i];
    printf("TMEM fill/read-back check: %s, mismatching doubles: %.0f\n", cudaGetErrorString(err), bad);
    run<0>("shared only", out, cyc, sms);
    run<1>("TMEM only", out, cyc, sms);
    run<2>("half TMEM + half shared", out, cyc, sms);
    return 0;
}
