// Micro-benchmark: how FP64 instructions share the issue port with other instructions on sm_100a.
// 16 warps per SM (4 per scheduler), each with 8 independent accumulator chains so that latency never limits.
//   mode 0: DADD only          mode 1: DADD + one IADD3 per DADD          mode 2: DADD + one LDS.64 per DADD
//   mode 3: IADD only          mode 4: DADD + two IADD per DADD
// If an FP64 warp-instruction occupies its scheduler for 2 cycles, mode 1 takes (2+1)/2 of mode 0; if the integer instruction
// issues in the FP64 pipe's shadow, mode 1 == mode 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o ubench_fp64_issue tools/ubench_fp64_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 1 << 14;

template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(double* out, long long* cyc, int salt) {
    __shared__ double sm[512];
    sm[threadIdx.x] = 1.0 + threadIdx.x * 1e-9;
    __syncthreads();
    double a[8];
    int q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = threadIdx.x * 1e-3 + u; q[u] = threadIdx.x + u + salt; }
    const double inc = 1.0 + salt * 1e-12;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE != 3) a[u] += inc;
            if (MODE == 1 || MODE == 3 || MODE == 4) q[u] = q[u] * 3 + it;          // IMAD
            if (MODE == 4) q[u] ^= (q[u] >> 3);                                    // + SHF/LOP
            if (MODE == 2) a[u] += 0.0 * sm[(threadIdx.x + u + it) & 511];          // LDS + DMUL + DADD -> counted below
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
    int z = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) { s += a[u]; z += q[u]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + z;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double* out, long long* cyc, int sms, double fp64_per_elem, double other_per_elem) {
    bench<MODE><<<sms, 512>>>(out, cyc, 1);
    cudaDeviceSynchronize();
    bench<MODE><<<sms, 512>>>(out, cyc, 2);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    const double elems = (double)kIters * 8 * 16;   // per SM: warp-level "elements"
    printf("%-34s %s cycles %lld -> %.3f cycles per element per scheduler (FP64 %.0f + other %.0f warp-instr per element)\n", name,
           cudaGetErrorString(e), c, (double)c / (elems / 4), fp64_per_elem, other_per_elem);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * 512);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    run<0>("DADD only", out, cyc, sms, 1, 0);
    run<3>("IMAD only", out, cyc, sms, 0, 1);
    run<1>("DADD + IMAD", out, cyc, sms, 1, 1);
    run<4>("DADD + IMAD + SHF + LOP", out, cyc, sms, 1, 3);
    run<2>("DADD + (LDS, DMUL, DADD)", out, cyc, sms, 3, 1);
    return 0;
}
