// lmi_device.cuh — LMIOracle / LMIOldOracle / LMI0Oracle with LDLTMgr on the device, one warp per problem
// (src/oracles/lmi_oracle.rs:24-44, lmi_old_oracle.rs, lmi0_oracle.rs:17-34; src/oracles/ldlt_mgr.rs:29-55, 98-111, 115-124).
// Shared by batch_stable.cu (EllStable space) and batch_ell64.cu (Ell space): the reference runs its LMI oracles under both
// (tests/lmi_tests.rs:174-225).  n <= 64 variables (lane L owns indices L and L+32), LMI size m <= 32.
#pragma once
#include "common.cuh"

namespace ellcuda {
namespace {

constexpr int kNMax = 64;

__device__ __forceinline__ int tri_index(int i, int j) {  // packed lower triangle, j <= i
    return i * (i + 1) / 2 + j;
}
__device__ __forceinline__ int upper_off(int r, int n) {  // packed upper triangle incl. diagonal: row r starts here
    return r * n - r * (r - 1) / 2;
}
__device__ __forceinline__ double bcast(double lo, double hi, int idx) {
    const double a = __shfl_sync(0xffffffffu, lo, idx & 31);
    const double b = __shfl_sync(0xffffffffu, hi, idx & 31);
    return idx < 32 ? a : b;
}

struct WarpMem {
    double* sx;    // n   xc (oracle input)
    double* sv;    // n   general scratch vector (products / gg)
    double* sw;    // n   w = inv(L) g
    double* sz;    // n   z
    double* sa;    // tri(m)  LMI elements A(i,j), j <= i
    double* st;    // m*m LDLT storage
    double* swit;  // m   witness
    double* qu;    // GENERAL only: packed upper triangle of the EllStable matrix
};

// ---------------------------------------------------------------------------------------------- LMI oracle
struct LmiShared {
    const double* fp;  // n × fstride packed lower triangles of F_k
    const double* bp;  // tri packed lower triangle of B
    int fstride, m, tri, nvar;
    int lmi0;          // LMI0Oracle (lmi0_oracle.rs:17-34): elements start from 0.0 and ADD F_k*x_k; gradient negated; no B
};

// returns true if a cut was produced: g0/g1 = this lane's gradient entries, ep = beta
__device__ bool lmi_assess(const LmiShared& L, const WarpMem& wm, int lane, double& g0, double& g1, double& ep) {
    const int m = L.m, tri = L.tri, n = L.nvar;
    // --- elements A(i,j) = B[i][j] - sum_k F_k[i][j]*x[k], k ascending (src/oracles/lmi_oracle.rs:28-34)
    // 5 elements per lane per pass: m = 16 (136 elements) is one pass with every lane busy.  Slots past the last element
    // work on a valid element (the last one) and are dropped at the store, so the inner loop carries no predicate.
    constexpr int U = 5;
    for (int e0 = lane; e0 < tri; e0 += 32 * U) {
        double a[U];
        int eu[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + 32 * u;
            eu[u] = e < tri ? e : tri - 1;
            a[u] = L.lmi0 ? 0.0 : L.bp[eu[u]];
        }
        if (!L.lmi0) {
#pragma unroll 2
            for (int k = 0; k < n; ++k) {
                const double xk = wm.sx[k];
                const double* fk = L.fp + k * L.fstride;
#pragma unroll
                for (int u = 0; u < U; ++u) a[u] -= fk[eu[u]] * xk;
            }
        } else {
#pragma unroll 2
            for (int k = 0; k < n; ++k) {
                const double xk = wm.sx[k];
                const double* fk = L.fp + k * L.fstride;
#pragma unroll
                for (int u = 0; u < U; ++u) a[u] += fk[eu[u]] * xk;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + 32 * u;
            if (e < tri) wm.sa[e] = a[u];
        }
    }
    __syncwarp();
    // --- LDLTMgr::factor (src/oracles/ldlt_mgr.rs:29-55)
    int pos = 0;
    if (m <= 16) {
        // Two rows at a time, one per half-warp, the second one step behind the first: step j of row i+1 needs T[j][i], which
        // step j of row i has just produced, and nothing later — so the m^2/2 serial steps (each with an IEEE division on its
        // critical path) become i+2 ticks per pair of rows, 72 instead of 120 for m = 16.  Every element still sees the
        // reference's operations in the reference's order; a row started speculatively after a failing one only writes
        // entries that nothing reads.  Lane = (half h, column c).
        const int h = lane >> 4, c = lane & 15;
        for (int i = 0; i < m && pos == 0; i += 2) {
            const int ih = i + h;                                   // this half's row
            const bool row_on = ih < m;
            double acc = 0.0;                                       // s for column c: sum_k T[ih][k]*T[k][c], k ascending
            double diag = row_on ? wm.sa[tri_index(ih, 0)] : 1.0;   // get_elem(ih, start)
            for (int t = 0; t <= i + 1; ++t) {
                if (t == i) {                                       // row i is complete: its pivot
                    const double dii = __shfl_sync(0xffffffffu, diag, i);
                    if (lane == 0) wm.st[i * m + i] = dii;
                    if (dii <= 0.0) { pos = i + 1; break; }
                    __syncwarp();
                }
                const int j = t - h;
                const bool step_on = row_on && j >= 0 && j < ih;
                const int jj = step_on ? j : 0;
                const double d = __shfl_sync(0xffffffffu, diag, (lane & 16) | jj);
                const double val = d / wm.st[jj * m + jj];
                if (step_on) {
                    if (c == 0) {
                        wm.st[j * m + ih] = d;                      // keep for later
                        wm.st[ih * m + j] = val;                    // L[ih][j]
                    }
                    if (c > j && c <= ih) {
                        const double tjc = (c == ih) ? d : wm.st[j * m + c];
                        acc += val * tjc;
                        if (c == j + 1) diag = wm.sa[tri_index(ih, c)] - acc;   // get_elem(ih, stop) - s
                    }
                }
                __syncwarp();
            }
            if (pos == 0 && i + 1 < m) {                            // the pivot of row i+1
                const double dii = __shfl_sync(0xffffffffu, diag, 16 | (i + 1));
                if (lane == 0) wm.st[(i + 1) * m + (i + 1)] = dii;
                if (dii <= 0.0) pos = i + 2;
            }
            __syncwarp();
        }
    } else {
        for (int i = 0; i < m; ++i) {                               // lane c follows column c of the current row
            double acc = 0.0;                                       // s for column `lane`: sum_k T[i][k]*T[k][lane], k ascending
            double diag = wm.sa[tri_index(i, 0)];                   // get_elem(i, start)
            for (int j = 0; j < i; ++j) {
                const double d = __shfl_sync(0xffffffffu, diag, j);
                const double val = d / wm.st[j * m + j];
                if (lane == 0) {
                    wm.st[j * m + i] = d;                           // keep for later
                    wm.st[i * m + j] = val;                         // L[i][j]
                }
                const int c = lane;
                if (c > j && c <= i) {
                    const double tjc = (c == i) ? d : wm.st[j * m + c];
                    acc += val * tjc;
                    if (c == j + 1) diag = wm.sa[tri_index(i, c)] - acc;   // get_elem(i, stop) - s
                }
            }
            const double dii = __shfl_sync(0xffffffffu, diag, i);
            if (lane == 0) wm.st[i * m + i] = dii;
            __syncwarp();
            if (dii <= 0.0) { pos = i + 1; break; }
        }
    }
    if (pos == 0) return false;                             // is_spd
    // --- witness (ldlt_mgr.rs:98-111), start = 0: wit[i-1] = -(sum_{k=i}^{pos-1} T[k][i-1]*wit[k]), k ascending.  Lane k keeps
    // wit[k]; the products of a row are formed by their lanes, the ordered sum runs over shared memory.
    const int mm = pos - 1;
    double wk = lane == mm ? 1.0 : 0.0;
    for (int i = mm; i >= 1; --i) {
        if (lane >= i && lane < pos) wm.swit[lane] = wm.st[lane * m + (i - 1)] * wk;   // swit doubles as the product buffer here
        __syncwarp();
        double s = 0.0;
#pragma unroll 4
        for (int k = i; k < pos; ++k) s += wm.swit[k];
        if (lane == i - 1) wk = -s;
        __syncwarp();
    }
    if (lane < pos) wm.swit[lane] = wk;
    __syncwarp();
    ep = -wm.st[mm * m + mm];
    // --- g[k] = sym_quad(F_k) (ldlt_mgr.rs:115-124): i outer, j inner, (wit[i]*F[i][j])*wit[j]
    double r0 = 0.0, r1 = 0.0;
    const int k0 = lane, k1 = lane + 32;
    const double* f0 = L.fp + (k0 < n ? k0 : 0) * L.fstride;
    const double* f1 = L.fp + (k1 < n ? k1 : 0) * L.fstride;
    for (int i = 0; i < pos; ++i) {
        const double wi = wm.swit[i];
        int e = tri_index(i, 0);
        int j = 0;
        for (; j + 1 <= i; j += 2, e += 2) {         // F[i][j], j <= i: row i of the packed lower triangle (two per trip)
            const double wj0 = wm.swit[j], wj1 = wm.swit[j + 1];
            const double a0 = f0[e], a1 = f1[e], b0 = f0[e + 1], b1 = f1[e + 1];
            r0 += wi * a0 * wj0;
            r1 += wi * a1 * wj0;
            r0 += wi * b0 * wj1;
            r1 += wi * b1 * wj1;
        }
        if (j <= i) {
            const double wj = wm.swit[j];
            r0 += wi * f0[e] * wj;
            r1 += wi * f1[e] * wj;
        }
        e = tri_index(i + 1, i);
        j = i + 1;
        for (; j + 1 < pos; j += 2) {                 // F[i][j] = F[j][i], j > i: walk down column i (two per trip)
            const int e2 = e + j + 1;
            const double wj0 = wm.swit[j], wj1 = wm.swit[j + 1];
            const double a0 = f0[e], a1 = f1[e], b0 = f0[e2], b1 = f1[e2];
            r0 += wi * a0 * wj0;
            r1 += wi * a1 * wj0;
            r0 += wi * b0 * wj1;
            r1 += wi * b1 * wj1;
            e = e2 + j + 2;
        }
        if (j < pos) {
            const double wj = wm.swit[j];
            r0 += wi * f0[e] * wj;
            r1 += wi * f1[e] * wj;
        }
    }
    if (L.lmi0) { r0 = -r0; r1 = -r1; }   // g[k] = -sym_quad(F_k), lmi0_oracle.rs:31
    g0 = k0 < n ? r0 : 0.0;
    g1 = k1 < n ? r1 : 0.0;
    return true;
}


// stage F_k (packed lower triangles, stride `fstride`) and B once per CTA; returns the number of doubles used (even)
__device__ __forceinline__ long long lmi_stage_tables(const BatchArgs& a, double* smem, int fstride, int lmi0) {
    const int n = a.n, m = a.aux[0], tri = m * (m + 1) / 2;
    double* fp = smem;
    double* bp = fp + (long long)n * fstride;
    for (int idx = threadIdx.x; idx < n * tri; idx += blockDim.x) {
        const int k = idx / tri, e = idx - k * tri;
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (tri_index(i + 1, 0) <= e) ++i;
        while (tri_index(i, 0) > e) --i;
        const int j = e - tri_index(i, 0);
        fp[k * fstride + e] = a.table[(long long)k * m * m + i * m + j];
    }
    for (int e = threadIdx.x; e < tri; e += blockDim.x) {
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (tri_index(i + 1, 0) <= e) ++i;
        while (tri_index(i, 0) > e) --i;
        const int j = e - tri_index(i, 0);
        bp[e] = lmi0 ? 0.0 : a.table[(long long)n * m * m + i * m + j];
    }
    __syncthreads();
    long long base = (long long)n * fstride + tri;
    return (base + 1) & ~1LL;
}

}  // namespace
}  // namespace ellcuda
