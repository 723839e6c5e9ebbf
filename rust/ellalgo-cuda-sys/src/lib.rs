//! Raw bindings to `include/ellcuda.h`.  NOT COMPILED IN THE BUILD IMAGE (no cargo/rustc there); kept in sync with
//! the header by hand — every item names the header line it mirrors.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const ELLCUDA_OK: c_int = 0;
pub const ELLCUDA_E_ARG: c_int = -1;
pub const ELLCUDA_E_CUDA: c_int = -2;

pub const ELL_KIND_ELL: i32 = 0;
pub const ELL_KIND_ELLSTABLE: i32 = 1;
pub const ELL_UPDATE_BIAS: i32 = 0;
pub const ELL_UPDATE_CENTRAL: i32 = 1;
pub const ELL_UPDATE_Q: i32 = 2;
pub const ELL_MODE_OPTIM: i32 = 0;
pub const ELL_MODE_FEAS: i32 = 1;
pub const ELL_MODE_OPTIM_Q: i32 = 2;
pub const ELL_MODE_BSEARCH: i32 = 3;
// built-in device oracles (include/ellcuda.h cites the reference file of each)
pub const ELL_ORACLE_EXAMPLE1: i32 = 0;
pub const ELL_ORACLE_EXAMPLE1_RR: i32 = 1;
pub const ELL_ORACLE_EXAMPLE4: i32 = 2;
pub const ELL_ORACLE_QUASICVX: i32 = 3;
pub const ELL_ORACLE_PROFIT: i32 = 4;
pub const ELL_ORACLE_PROFIT_RB: i32 = 5;
pub const ELL_ORACLE_PROFIT_Q: i32 = 6;
pub const ELL_ORACLE_LOWPASS: i32 = 7;
pub const ELL_ORACLE_BOWL: i32 = 8;
pub const ELL_ORACLE_LMI: i32 = 9;
pub const ELL_ORACLE_EXAMPLE2: i32 = 10;
pub const ELL_ORACLE_EXAMPLE3: i32 = 11;
pub const ELL_ORACLE_HALFPLANE: i32 = 12;
pub const ELL_ORACLE_ALWAYS_CUT: i32 = 13;
pub const ELL_ORACLE_OPTIM_BOX: i32 = 14;
pub const ELL_ORACLE_LMI_LINOPT: i32 = 15;
pub const ELL_ORACLE_LMI_OLD: i32 = 16;
pub const ELL_ORACLE_LMI0: i32 = 17;

#[repr(C)]
pub struct ellcuda_ctx {
    _private: [u8; 0],
}
pub type ell_handle = *mut ellcuda_ctx;

#[repr(C)]
pub struct ell_create_desc {
    pub kind: i32,
    pub device: i32,
    pub n: i64,
    pub kappa: f64,
    pub mq: *const f64,
    pub diag: *const f64,
    pub xc: *const f64,
    pub no_defer_trick: i32,
    pub use_parallel_cut: i32,
    pub row_begin: i64,
    pub row_end: i64,
    pub reduce_mode: i32,
    pub storage: i32, // ELL_STORAGE_FULL = 0 (bit-identical), ELL_STORAGE_SYMMETRIC = 1 (half the HBM traffic, 1e-9)
    pub stream: *mut c_void,
    pub shard_rank: i32,  // symmetric storage only: owner of the column-block pairs p = shard_rank (mod shard_count)
    pub shard_count: i32, // 0 or 1 = not sharded
}

#[repr(C)]
pub struct ell_batch_desc {
    pub kind: i32,
    pub oracle_id: i32,
    pub mode: i32,
    pub n: i32,
    pub count: i64,
    pub max_iters: i64,
    pub tolerance: f64,
    pub no_defer_trick: i32,
    pub use_parallel_cut: i32,
    pub kappa0: *const f64,
    pub kappa0_stride: i64,
    pub xc0: *const f64,
    pub xc0_stride: i64,
    pub diag0: *const f64,
    pub diag0_stride: i64,
    pub mq0: *const f64,
    pub mq0_stride: i64,
    pub gamma0: *const f64,
    pub gamma0_stride: i64,
    pub params: *const f64,
    pub params_stride: i64,
    pub params_width: i32,
    pub _pad0: i32,
    pub table: *const f64,
    pub table_len: i64,
    pub aux: [i32; 8],
    pub device: i32,
    pub on_device: i32,
    pub stream: *mut c_void,
}

#[repr(C)]
pub struct ell_batch_result {
    pub xbest: *mut f64,
    pub has_xbest: *mut u8,
    pub num_iters: *mut i64,
    pub exit_status: *mut i32,
    pub gamma: *mut f64,
    pub tsq: *mut f64,
    pub xc: *mut f64,
    pub kappa: *mut f64,
}

extern "C" {
    pub fn ellcuda_batch_solve(desc: *const ell_batch_desc, res: *const ell_batch_result) -> c_int;
    pub fn ellcuda_lowpass_table(n: i32, out: *mut f64) -> c_int;
    pub fn ellcuda_create(desc: *const ell_create_desc, out: *mut ell_handle) -> c_int;
    pub fn ellcuda_create_sharded(desc: *const ell_create_desc, ndev: i32, devs: *const i32, out: *mut ell_handle) -> c_int;
    pub fn ellcuda_group_size(h: ell_handle) -> i32;
    pub fn ellcuda_batch_solve_multi(desc: *const ell_batch_desc, res: *const ell_batch_result, ndev: i32, devs: *const i32) -> c_int;
    pub fn ellcuda_seq_sum(h: ell_handle, v_dev: *const f64, n: i64, out_host: *mut f64) -> c_int;
    pub fn ellcuda_set_no_defer_trick(h: ell_handle, on: i32) -> c_int;
    pub fn ellcuda_set_use_parallel_cut(h: ell_handle, on: i32) -> c_int;
    pub fn ellcuda_clone(h: ell_handle, out: *mut ell_handle) -> c_int;
    pub fn ellcuda_destroy(h: ell_handle);
    pub fn ellcuda_xc(h: ell_handle, out_n: *mut f64) -> c_int;
    pub fn ellcuda_set_xc(h: ell_handle, x_n: *const f64) -> c_int;
    pub fn ellcuda_tsq(h: ell_handle, out: *mut f64) -> c_int;
    pub fn ellcuda_kappa(h: ell_handle, out: *mut f64) -> c_int;
    pub fn ellcuda_get_mq(h: ell_handle, out_rows_by_n: *mut f64) -> c_int;
    pub fn ellcuda_update(
        h: ell_handle,
        which: i32,
        grad_host: *const f64,
        beta0: f64,
        has_beta1: i32,
        beta1: f64,
        out_status: *mut i32,
    ) -> c_int;
    pub fn ellcuda_last_error() -> *const c_char;
}
