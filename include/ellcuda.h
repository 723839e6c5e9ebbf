/* ellcuda.h — C ABI of libellcuda.so: the B200 (sm_100a) ellipsoid-update path.
 *
 * This header is the drop-in boundary.  The reference crate (luk036/ellalgo-rs v0.1.7)
 * has no FFI of its own; its extension surface is the trait set
 *   SearchSpace / UpdateByCutChoice            src/cutting_plane.rs:104-182
 *   OracleFeas / OracleOptim / OracleOptimQ    src/cutting_plane.rs:119-147
 * and the free drivers cutting_plane_feas/_optim/_optim_q (src/cutting_plane.rs:205-374).
 * Every entry point below names the reference item a Rust shim would forward to it
 * (INTEGRATION.md shows that shim).  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *   - return 0 = ok; negative = infrastructure error (ELLCUDA_E_*), text via
 *     ellcuda_last_error().  Algorithmic outcomes (CutStatus) travel only through
 *     out-parameters.  The reference panics on shape errors (src/arr.rs:427-429); the
 *     Rust shim turns ELLCUDA_E_ARG into panic!.
 *   - all floating point is IEEE binary64, never contracted to FMA, sums in the
 *     reference's left-to-right order (src/arr.rs:431-436, 446-450).
 *   - there is NO CPU fallback: every compute call fails with ELLCUDA_E_CUDA when no
 *     sm_100 device is usable.
 *   - a handle is not thread-safe (the reference's `&mut self`); distinct handles are
 *     independent.
 */
#ifndef ELLCUDA_H
#define ELLCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes -------------------------------------------------------------------- */
#define ELLCUDA_OK       0
#define ELLCUDA_E_ARG   (-1) /* bad argument / shape (reference: assert! panic)            */
#define ELLCUDA_E_CUDA  (-2) /* CUDA runtime error, or no usable sm_100 device             */
#define ELLCUDA_E_COMM  (-3) /* multi-GPU exchange error                                   */
#define ELLCUDA_E_OOM   (-4) /* device allocation failed                                   */

/* ---- CutStatus, enum order of src/cutting_plane.rs:32-37 ------------------------------ */
#define ELL_SUCCESS   0
#define ELL_NOSOLN    1
#define ELL_NOEFFECT  2
#define ELL_UNKNOWN   3

/* ---- why a cutting_plane_* run stopped.  The reference returns only (Option<x>, niter);
 *      north_star asks for the termination status too, so the batch API exports it.
 *      1,2 reuse the CutStatus codes.                                                      */
#define ELL_EXIT_NOSOLN    1 /* update returned NoSoln   (src/cutting_plane.rs:222,308,358) */
#define ELL_EXIT_NOEFFECT  2 /* update returned NoEffect (…:222,308; optim_q :361-364)      */
#define ELL_EXIT_TOL       4 /* Success but tsq < tolerance (…:222,308,369)                 */
#define ELL_EXIT_MAXIT     5 /* loop ran out (…:226,312,373)                                */
#define ELL_EXIT_FEASIBLE  6 /* cutting_plane_feas: oracle returned None (…:217-219)        */

/* ---- search-space kind ---------------------------------------------------------------- */
#define ELL_KIND_ELL        0 /* src/ell.rs        */
#define ELL_KIND_ELLSTABLE  1 /* src/ell_stable.rs */

/* ---- storage of a single-instance Ell (ell_create_desc.storage) ------------------------- */
#define ELL_STORAGE_FULL       0
#define ELL_STORAGE_SYMMETRIC  1

/* ---- which SearchSpace method — src/cutting_plane.rs:161-180 -------------------------- */
#define ELL_UPDATE_BIAS     0 /* update_bias_cut    */
#define ELL_UPDATE_CENTRAL  1 /* update_central_cut */
#define ELL_UPDATE_Q        2 /* update_q           */

/* ---- driver ---------------------------------------------------------------------------- */
#define ELL_MODE_OPTIM    0 /* cutting_plane_optim   src/cutting_plane.rs:286-313 */
#define ELL_MODE_FEAS     1 /* cutting_plane_feas    src/cutting_plane.rs:205-227 */
#define ELL_MODE_OPTIM_Q  2 /* cutting_plane_optim_q src/cutting_plane.rs:331-374 */
#define ELL_MODE_BSEARCH  3 /* bsearch over BSearchAdaptor<Oracle, Space>  src/cutting_plane.rs:376-466:
                               gamma0[i] = lower end, params[i] = {upper end, outer tolerance, outer max_iters};
                               max_iters / tolerance are the adaptor's (inner) Options.  Results: num_iters = outer
                               iterations, has_xbest = `upper != u_orig` (the bool bsearch returns), gamma = final upper,
                               xbest = the adaptor's space.xc() (last feasible point).  Oracle: ELL_ORACLE_EXAMPLE3.   */

/* ---- built-in device oracles (the reference's own oracle structs, restated on device) -- */
#define ELL_ORACLE_EXAMPLE1     0 /* src/example1.rs:8-30            optim n=2  no params   */
#define ELL_ORACLE_EXAMPLE1_RR  1 /* src/example1_rr.rs:13-53        optim n=2  no params   */
#define ELL_ORACLE_EXAMPLE4     2 /* src/example4.rs:10-57           optim n=2  no params   */
#define ELL_ORACLE_QUASICVX     3 /* src/quasicvx.rs:11-51           optim n=2  no params   */
#define ELL_ORACLE_PROFIT       4 /* src/oracles/profit_oracle.rs:42-114  optim n=2
                                     params[7] = p, A, k, a0, a1, v0, v1                    */
#define ELL_ORACLE_PROFIT_RB    5 /* …:116-161  params[12] = the 7 above + e1..e5           */
#define ELL_ORACLE_PROFIT_Q     6 /* …:163-198  optim_q, params[7]                          */
#define ELL_ORACLE_LOWPASS      7 /* src/oracles/lowpass_oracle.rs:7-150  optim, n taps, n <= 224
                                     (tests/stress_tests.rs:6-14 uses 128)
                                     params[5] = wpass, wstop, lp_sq, up_sq, sp_sq
                                     table = spectrum, aux[0] = mdim rows × n (see
                                     ellcuda_lowpass_table)                                 */
#define ELL_ORACLE_BOWL         8 /* tests/integration_test.rs:90-105 optim, n <= 224
                                     params[n] = target.  Batched `Ell` keeps a problem's matrix
                                     on the SM: one warp with the rows in registers (n <= 32) or
                                     tensor memory (n <= 64), one CTA with the packed lower
                                     triangle in shared memory (n <= 224); larger n: the
                                     single-instance handle below                           */
#define ELL_ORACLE_LMI          9 /* src/oracles/lmi_oracle.rs:5-45 + ldlt_mgr.rs  feas
                                     table = F_0..F_{n-1} (each m×m row-major) then B (m×m);
                                     aux[0] = m                                             */
#define ELL_ORACLE_EXAMPLE2    10 /* tests/example2_tests.rs:12-48   feas  n=2              */
#define ELL_ORACLE_EXAMPLE3    11 /* src/example3.rs:10-60           feas  n=2 params[1]=target */
#define ELL_ORACLE_HALFPLANE   12 /* tests/cutting_plane_tests.rs:12-28  feas n=2           */
#define ELL_ORACLE_ALWAYS_CUT  13 /* tests/cutting_plane_tests.rs:33-42, 91-99  feas|optim  */
#define ELL_ORACLE_OPTIM_BOX   14 /* tests/cutting_plane_tests.rs:47-73  optim n=2          */
#define ELL_ORACLE_LMI_LINOPT  15 /* tests/lmi_tests.rs:121-172 MyLmiOracle: min c·x over L
                                     LMI blocks, round-robin.  table = c[n], then per block
                                     F_0..F_{n-1}, B;  aux[0] = L, aux[1..L] = block sizes
                                     (device: n = 3, L <= 3, block size <= 4)               */
#define ELL_ORACLE_LMI_OLD     16 /* src/oracles/lmi_old_oracle.rs:5-50  feas; table and aux as ELL_ORACLE_LMI.  The
                                     reference forms all of A = B - sum_k F_k x_k before factorizing; each element sees
                                     the same operations in the same order, so results equal ELL_ORACLE_LMI's bit for bit
                                     (the device evaluates only the elements the factorization asks for)             */
#define ELL_ORACLE_LMI0        17 /* src/oracles/lmi0_oracle.rs:4-35  feas: sum_k F_k x_k > 0, cut (-sym_quad(F_k),
                                     SingleCut(ep)).  table = F_0..F_{n-1} only; aux[0] = m                          */
#define ELL_ORACLE_COUNT       18

/* ======================================================================================= *
 *  Batched mode: `count` independent problems, the whole cutting-plane loop on-device.
 *  Replaces, per problem i:
 *      let mut space = Ell::new_with_matrix(kappa0[i], mq0 | diag(diag0[i]) | I, xc0[i]);
 *      let mut omega = <Oracle>::new(params[i], table);
 *      let (xbest, num_iters) = cutting_plane_{optim,feas,optim_q}(&mut omega, &mut space,
 *                                   &mut gamma[i], &Options{max_iters, tolerance});
 *  (src/ell.rs:31-78, src/ell_stable.rs:18-35, src/cutting_plane.rs:50-100, 205-374).
 *
 *  Per-instance arrays are SoA with an explicit stride in doubles between instances;
 *  stride 0 means one value shared by every instance.
 * ======================================================================================= */
typedef struct ell_batch_desc {
    int32_t kind;             /* ELL_KIND_*                                               */
    int32_t oracle_id;        /* ELL_ORACLE_*                                             */
    int32_t mode;             /* ELL_MODE_*                                               */
    int32_t n;                /* dimension of xc                                          */
    int64_t count;            /* number of independent problems                           */
    int64_t max_iters;        /* Options::max_iters  (default 2000, …:97-99)              */
    double  tolerance;        /* Options::tolerance  (default 1e-20)                      */
    int32_t no_defer_trick;   /* Ell::no_defer_trick  src/ell.rs:10,132-135               */
    int32_t use_parallel_cut; /* EllCalc::use_parallel_cut src/ell_calc.rs:630 (default 1) */

    const double* kappa0;  int64_t kappa0_stride;  /* 1 value / instance                  */
    const double* xc0;     int64_t xc0_stride;     /* n values / instance                 */
    const double* diag0;   int64_t diag0_stride;   /* n values; NULL ⇒ identity           */
    const double* mq0;     int64_t mq0_stride;     /* n·n values row-major; NULL ⇒ diag0  */
    const double* gamma0;  int64_t gamma0_stride;  /* 1 value; ignored in FEAS mode       */
    const double* params;  int64_t params_stride;  /* params_width values / instance      */
    int32_t params_width;
    int32_t _pad0;
    const double* table;   int64_t table_len;      /* shared read-only table (doubles)    */
    int32_t aux[8];           /* oracle-specific integers, see ELL_ORACLE_*               */

    int32_t device;           /* CUDA device ordinal                                      */
    int32_t on_device;        /* 0: every pointer is host memory (copied in/out inside the
                                    call); 1: every pointer is device memory on `device`   */
    void*   stream;           /* cudaStream_t to run on; NULL ⇒ the library's own stream (to name
                                 the legacy default stream pass cudaStreamLegacy, (void*)1)      */
} ell_batch_desc;

typedef struct ell_batch_result { /* caller-allocated; optional members may be NULL       */
    double*  xbest;       /* count·n  Option<xbest> payload (unchanged where has_xbest=0)  */
    uint8_t* has_xbest;   /* count    1 = Some                                             */
    int64_t* num_iters;   /* count    second tuple member of the driver's return value     */
    int32_t* exit_status; /* count    ELL_EXIT_*                                           */
    double*  gamma;       /* count    final gamma (optim modes), optional                  */
    double*  tsq;         /* count    final space.tsq(), optional                          */
    double*  xc;          /* count·n  final space.xc(),  optional                          */
    double*  kappa;       /* count    final kappa,       optional                          */
} ell_batch_result;

/* Runs the batch on desc->device.  Synchronous when on_device==0; when on_device==1 the
 * work is enqueued on desc->stream and the call returns without synchronising.          */
int ellcuda_batch_solve(const ell_batch_desc* desc, const ell_batch_result* res);

/* The same batch spread over `ndev` devices of this process (SURVEY.md §8b `ndev, devs`; SURVEY.md §8e: independent
 * problems, contiguous blocks, no data-path collective).  Host buffers only (desc->on_device must be 0; desc->device and
 * desc->stream are ignored): device r solves problems [r·count/ndev, (r+1)·count/ndev) on its own stream from its own host
 * thread, and the call returns when every block's results are back in the caller's arrays.                            */
int ellcuda_batch_solve_multi(const ell_batch_desc* desc, const ell_batch_result* res, int32_t ndev, const int32_t* devs);

/* Number of kernels the last ellcuda_batch_solve(_multi) on this thread launched (bench counter). */
int64_t ellcuda_last_launch_count(void);

/* LowpassOracle::new's spectrum table (src/oracles/lowpass_oracle.rs:24-33 with
 * linspace src/arr.rs:506-520): out[i*n+j], i < 15·n.  Host-side, glibc cos, so the table
 * the device reads is bit-identical to the one the Rust constructor builds.             */
int ellcuda_lowpass_table(int32_t n, double* out_mdim_by_n);
/* nwpass / nwstop of LowpassOracle::new (…:35-36). */
int32_t ellcuda_lowpass_band_edge(int32_t n, double w);

/* ======================================================================================= *
 *  Single-instance mode (large n): an `Ell` whose matrix lives in HBM.
 *  The handle stands for one `Ell` value (src/ell.rs:9-16).  With row sharding
 *  (row_begin,row_end) it stands for one rank's slab of rows; ranks exchange the slab
 *  results of `grad_t` between ellcuda_update_begin and ellcuda_update_finish.
 * ======================================================================================= */
typedef struct ellcuda_ctx* ell_handle;

typedef struct ell_create_desc {
    int32_t kind;              /* ELL_KIND_ELL: any n, matrix streamed from HBM, optionally row-sharded.
                                  ELL_KIND_ELLSTABLE (src/ell_stable.rs): n <= 2048, whole matrix, `mq` = the
                                  factor as the reference holds it (diagonal + strict upper triangle are
                                  read; need not be symmetric); its triangular recurrences neither stream
                                  nor shard (one warp for n <= 64, one CTA above; serial in the reference's
                                  summation order) — throughput comes from ellcuda_batch_solve.
                                  Cost of one update_core on a B200: ~25 us at n = 100, ~0.7 ms at n = 512,
                                  ~2.7 ms at n = 1024, ~11 ms at n = 2048: the back substitution is one chain of
                                  n^2/2 dependent FP64 subtractions (10 clocks each) because the reference sums
                                  each row left to right starting from the entry that finished last.  The
                                  batched EllStable kernel runs 7.9e7 such updates/s at n = 64 (DIAG start)     */
    int32_t device;
    int64_t n;
    double  kappa;             /* Ell::new_with_scalar(kappa,…) / new_with_matrix(kappa,…) */
    const double* mq;          /* host, n·n row-major; NULL ⇒ diag.  Need not be symmetric: as in the reference
                                  (src/ell.rs:31-41, 102, 117-128) the rows are used as given until the first
                                  successful cut rewrites the upper triangle from the lower one; such a handle
                                  keeps a second copy of the matrix until then and its ellcuda_update_finish
                                  synchronises to learn the status                                           */
    const double* diag;        /* host, n; NULL ⇒ identity   (Ell::new: kappa=1, diag=val) */
    const double* xc;          /* host, n                                                 */
    int32_t no_defer_trick;
    int32_t use_parallel_cut;
    int64_t row_begin, row_end;/* this handle's slab of rows; 0,n for a whole matrix       */
    int32_t reduce_mode;       /* 0: reference order (bit-identical grad_t and omega);
                                  k>0: each row sum split into k ordered segments and omega
                                  reduced blockwise — deterministic, ≤1e-13 rel. of mode 0 */
    int32_t storage;           /* ELL_STORAGE_FULL (0): the full mirrored matrix as the reference holds it; results
                                  bit-identical in reduce_mode 0.  ELL_STORAGE_SYMMETRIC (1): only the lower
                                  block-triangle is kept and streamed — half the HBM traffic per update; the sums of
                                  M·g are regrouped, so results are held to 1e-9 relative, not bit-identical.  `mq`
                                  must be symmetric; shards by shard_rank / shard_count below (or through
                                  ellcuda_create_sharded), not by rows                                            */
    void*   stream;            /* cudaStream_t; NULL ⇒ own stream, (void*)1 = cudaStreamLegacy */
    int32_t shard_rank;        /* ELL_STORAGE_SYMMETRIC only: this handle keeps current the tiles of the column-block   */
    int32_t shard_count;       /* pairs p ≡ shard_rank (mod shard_count) and the ranks REDUCE their partial grad_t over
                                  NVLink after ellcuda_exchange_attach (0 or 1 ⇒ not sharded).  Full storage shards by
                                  row_begin / row_end instead (a gather).  ellcuda_create_sharded sets these itself      */
} ell_create_desc;

int  ellcuda_create(const ell_create_desc* desc, ell_handle* out); /* Ell::new* src/ell.rs:31-78 */
/* The same Ell with its rows sharded over `ndev` devices of THIS process (SURVEY.md §8b `ndev, devs`): one slab per device,
 * exchange blocks attached by pointer (peer access over NVLink), no launcher / torch / NCCL needed.  desc->device, row_begin,
 * row_end and stream are ignored (each slab runs on its own stream); desc->reduce_mode applies to every slab (0 keeps the
 * result bit-identical to the reference, k = ndev row segments keeps all SMs streaming).  The handle is used like any other:
 * ellcuda_update (host gradient), update_begin/finish (the device gradient is copied to every slab's device and must be
 * complete when begin is called), xc / set_xc / tsq / kappa / get_mq (all n rows) / dot_mv (all n entries) / clone / destroy.
 * `devs` may name a device more than once: the slabs of one GPU then share one stream, so that no kernel ever waits for a
 * later launch on the same GPU.                                                                                           */
int  ellcuda_create_sharded(const ell_create_desc* desc, int32_t ndev, const int32_t* devs, ell_handle* out);
int32_t ellcuda_group_size(ell_handle h);                          /* devices behind the handle (1 for a plain one)  */
int  ellcuda_clone(ell_handle h, ell_handle* out);                 /* #[derive(Clone)] src/ell.rs:8 */
void ellcuda_destroy(ell_handle h);                                /* Drop                     */

int ellcuda_xc(ell_handle h, double* out_n);              /* SearchSpace::xc   src/ell.rs:144-146 */
int ellcuda_set_xc(ell_handle h, const double* x_n);      /* SearchSpace::set_xc      …:177-179   */
int ellcuda_tsq(ell_handle h, double* out);               /* SearchSpace::tsq         …:149-151   */
int ellcuda_kappa(ell_handle h, double* out);             /* pub field Ell::kappa     …:13        */
/* The two public knobs the reference lets a caller change after construction:
 *   Ell::no_defer_trick (pub field, src/ell.rs:10; written after new() at :342-344) — takes effect from the next update,
 *   exactly as the field does (src/ell.rs:132-135 reads it at the end of update_core); Ell handles only;
 *   EllCalc::use_parallel_cut (pub field of the helper, src/ell_calc.rs:630) — read by the next calc_* call.          */
int ellcuda_set_no_defer_trick(ell_handle h, int32_t on);
int ellcuda_set_use_parallel_cut(ell_handle h, int32_t on);
int ellcuda_get_mq(ell_handle h, double* out_rows_by_n);  /* pub field Ell::mq (this slab's rows,
                                                             pending rank-1 applied first)        */

/* Arr::dot_mv (src/arr.rs:426-442) on the handle's matrix: y = mq·x with the reference's left-to-right sums (mq is
 * symmetric).  x_dev: n doubles, y_dev: row_end-row_begin doubles, both device memory; enqueued on the handle's stream.
 * The streaming GEMV that power_iteration (src/power_iteration.rs) and conjugate_gradient
 * (src/conjugate_gradient.rs) share with Ell::update_core.  Ell handles only.                                       */
int ellcuda_dot_mv(ell_handle h, const double* x_dev, double* y_dev);
/* sum of v_dev[0..n) strictly left to right from 0.0 on the handle's device and stream; synchronous, result on the host.  With
 * the elementwise product formed by the caller this is Arr::dot (src/arr.rs:443-451) and the norms of
 * src/power_iteration.rs:22-25 for vectors that stay on the device between two ellcuda_dot_mv calls.                  */
int ellcuda_seq_sum(ell_handle h, const double* v_dev, int64_t n, double* out_host);

/* SearchSpace::update_{bias_cut,central_cut,q} with SingleCut(beta0) [has_beta1 = -1],
 * ParallelCut(beta0, None) [0] or ParallelCut(beta0, Some(beta1)) [1]
 * (src/ell.rs:153-175, 182-229 → update_core :97-137).  `grad` is host memory, n values.
 * Whole-matrix handles only.  Synchronous: *out_status is valid on return.               */
int ellcuda_update(ell_handle h, int32_t which, const double* grad_host,
                   double beta0, int32_t has_beta1, double beta1, int32_t* out_status);

/* Same, split for sharded / device-resident use.  `grad_dev` is device memory (n values) and must stay valid and
 * unchanged until the matching ellcuda_update_finish has run on the stream (both kernels read it; when n is a
 * multiple of 64 and the pointer is 16-byte aligned it is read in place, otherwise a padded copy is made).
 *   begin : applies the previous cut's deferred rank-1 to this slab and, in the same sweep
 *           over M, forms this slab's part of grad_t = M·grad   (ell.rs:102, 117-128)
 *   (exchange: every rank needs all n entries of grad_t — allgather of
 *    ellcuda_grad_t_ptr(h)[row_begin..row_end] in place; nothing to do for a whole matrix)
 *   finish: omega, tsq, EllCalc, xc, kappa (ell.rs:103-115,130-135), identically on every
 *           rank; the CutStatus is left in device memory (ellcuda_status_ptr) so that the
 *           sequence can be captured in a CUDA graph.                                      */
int ellcuda_update_begin(ell_handle h, int32_t which, const double* grad_dev,
                         double beta0, int32_t has_beta1, double beta1);
int ellcuda_update_finish(ell_handle h);
double*  ellcuda_grad_t_ptr(ell_handle h);   /* device, n doubles: grad_t of handles without peer exchange  */

/* Peer exchange over NVLink without NCCL (one process per GPU).  Every rank exports its exchange block
 * (ELLCUDA_IPC_HANDLE_BYTES opaque bytes, a CUDA IPC handle), the P handles are gathered by any means
 * (e.g. torch.distributed.all_gather_object) and attached; from then on ellcuda_update_begin stores this rank's
 * slab of grad_t directly into every rank's block from the kernel that produces it, and ellcuda_update_finish
 * waits (on the device) until all P slabs have arrived — no collective call between begin and finish.
 * attach_ptrs takes block pointers already valid in this process (ellcuda_exchange_block_ptr of handles living in the
 * same process; used to emulate ranks on one GPU, all on ONE stream, begin for every rank before any finish).     */
#define ELLCUDA_IPC_HANDLE_BYTES 64
int   ellcuda_exchange_export(ell_handle h, void* handle_out);
int   ellcuda_exchange_attach(ell_handle h, int32_t nranks, int32_t rank, const void* handles /* nranks x 64 B; one rank per GPU:
                                kernels of different ranks wait for each other's stores and must be able to run at the same time */);
int   ellcuda_exchange_attach_ptrs(ell_handle h, int32_t nranks, int32_t rank, void* const* blocks);
void* ellcuda_exchange_block_ptr(ell_handle h);
int32_t* ellcuda_status_ptr(ell_handle h);   /* device, CutStatus of the last finish       */
int ellcuda_last_status(ell_handle h, int32_t* out_status); /* synchronises, reads status   */
int ellcuda_set_stream(ell_handle h, void* stream);
int ellcuda_sync(ell_handle h);
/* duration in ms of the most recent M sweep on this handle (CUDA events on its stream);
 * negative if none was timed.  ellcuda_set_timing(h,1) enables event recording.           */
int ellcuda_set_timing(ell_handle h, int32_t on);
float ellcuda_last_sweep_ms(ell_handle h);
int64_t ellcuda_handle_launch_count(ell_handle h);

/* Diagnostic: y[i] = exp(x[i]) (which=0) / ln(x[i]) (which=1) computed with the device routines the built-in
 * oracles use; they reproduce glibc's f64 exp/ln (what Rust's f64::exp / f64::ln call, profit_oracle.rs:54-113)
 * bit for bit, and this entry point lets a test check that against the host libm.  Host buffers.            */
int ellcuda_math_probe(int32_t which, const double* x_host, double* y_host, int64_t n, int32_t device);

/* ---- misc -------------------------------------------------------------------------------- */
const char* ellcuda_last_error(void);   /* thread-local, never NULL                          */
int ellcuda_version(void);              /* 100*major + minor                                 */
int ellcuda_device_count(void);         /* usable sm_100 devices (0 ⇒ every compute call fails) */

#ifdef __cplusplus
}
#endif
#endif /* ELLCUDA_H */
