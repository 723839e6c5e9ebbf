//! `EllCuda` — drop-in for `ellalgo_rs::ell::Ell` (src/ell.rs) in the unchanged drivers
//! `cutting_plane_feas/_optim/_optim_q` and `BSearchAdaptor` (src/cutting_plane.rs:205-466).
//!
//! NOT COMPILED IN THE BUILD IMAGE (no cargo/rustc there).  The same calls are exercised by
//! `tests/cpp/test_host_mirror.cpp` (C++) and `tests/test_stream_gpu.py` (ctypes) against the identical symbols.
//!
//! ```ignore
//! use ellalgo_rs::cutting_plane::{cutting_plane_optim, Options};
//! let mut ellip = EllCuda::new_with_scalar(10.0, Arr::from(vec![0.0, 0.0]));     // was: Ell::new_with_scalar
//! let (xbest, num_iters) = cutting_plane_optim(&mut oracle, &mut ellip, &mut gamma, &Options::default());
//! ```
use ellalgo_cuda_sys as sys;
use ellalgo_rs::arr::Arr;
use ellalgo_rs::cutting_plane::{CutStatus, ParallelCut, SearchSpace, SingleCut, UpdateByCutChoice};
use std::ffi::CStr;

/// KIND = 0: `Ell` (src/ell.rs, any n, matrix in HBM); KIND = 1: `EllStable` (src/ell_stable.rs, n <= 2048).
pub struct EllCudaT<const KIND: i32> {
    h: sys::ell_handle,
    n: usize,
}
pub type EllCuda = EllCudaT<{ sys::ELL_KIND_ELL }>;
pub type EllStableCuda = EllCudaT<{ sys::ELL_KIND_ELLSTABLE }>;

// The handle owns device memory and one stream; `&mut self` serialises use exactly like the reference's plain data.
unsafe impl<const KIND: i32> Send for EllCudaT<KIND> {}

fn check(rc: i32) {
    if rc != sys::ELLCUDA_OK {
        let msg = unsafe { CStr::from_ptr(sys::ellcuda_last_error()) }.to_string_lossy().into_owned();
        // the reference panics on shape errors (src/arr.rs:427-429); infrastructure errors have no reference analogue
        panic!("libellcuda error {rc}: {msg}");
    }
}

impl<const KIND: i32> EllCudaT<KIND> {
    fn create(kappa: f64, mq: Option<&Arr>, diag: Option<&Arr>, xc: &Arr) -> Self {
        Self::create_on(kappa, mq, diag, xc, &[])
    }
    /// `devices` non-empty: the rows of M are sharded over these GPUs of this process (`ellcuda_create_sharded`) — the
    /// multi-GPU form of SURVEY §8b; the value is then used exactly like a single-GPU one.
    fn create_on(kappa: f64, mq: Option<&Arr>, diag: Option<&Arr>, xc: &Arr, devices: &[i32]) -> Self {
        let n = xc.len();
        let desc = sys::ell_create_desc {
            kind: KIND,
            device: 0,
            n: n as i64,
            kappa,
            mq: mq.map_or(std::ptr::null(), |m| m.data().as_ptr()),
            diag: diag.map_or(std::ptr::null(), |d| d.data().as_ptr()),
            xc: xc.data().as_ptr(),
            no_defer_trick: 0,
            use_parallel_cut: 1,
            row_begin: 0,
            row_end: n as i64,
            reduce_mode: 0,
            storage: 0,
            stream: std::ptr::null_mut(),
            shard_rank: 0,
            shard_count: 0,
        };
        let mut h: sys::ell_handle = std::ptr::null_mut();
        if devices.is_empty() {
            check(unsafe { sys::ellcuda_create(&desc, &mut h) });
        } else {
            check(unsafe { sys::ellcuda_create_sharded(&desc, devices.len() as i32, devices.as_ptr(), &mut h) });
        }
        EllCudaT { h, n }
    }
    /// `Ell::new_with_scalar` with M row-sharded over `devices` (large n)
    pub fn new_with_scalar_on(val: f64, xc: Arr, devices: &[i32]) -> Self {
        Self::create_on(val, None, None, &xc, devices)
    }
    /// `Ell::new_with_matrix` with M row-sharded over `devices`
    pub fn new_with_matrix_on(kappa: f64, mq: Arr, xc: Arr, devices: &[i32]) -> Self {
        assert_eq!(mq.size(), xc.len() * xc.len());
        Self::create_on(kappa, Some(&mq), None, &xc, devices)
    }
    /// the reference's `pub no_defer_trick` field is assigned after construction (src/ell.rs:10, 342-344):
    /// `ell.no_defer_trick = true;` becomes `ell.set_no_defer_trick(true);`
    pub fn set_no_defer_trick(&mut self, on: bool) {
        check(unsafe { sys::ellcuda_set_no_defer_trick(self.h, on as i32) });
    }
    /// `helper.use_parallel_cut` (src/ell_calc.rs:630)
    pub fn set_use_parallel_cut(&mut self, on: bool) {
        check(unsafe { sys::ellcuda_set_use_parallel_cut(self.h, on as i32) });
    }
    /// src/ell.rs:31-41 / src/ell_stable.rs:18-27
    pub fn new_with_matrix(kappa: f64, mq: Arr, xc: Arr) -> Self {
        assert_eq!(mq.size(), xc.len() * xc.len());
        Self::create(kappa, Some(&mq), None, &xc)
    }
    /// src/ell.rs:54-56
    pub fn new(val: Arr, xc: Arr) -> Self {
        Self::create(1.0, None, Some(&val), &xc)
    }
    /// src/ell.rs:65-67
    pub fn new_with_scalar(val: f64, xc: Arr) -> Self {
        Self::create(val, None, None, &xc)
    }
    /// src/ell.rs:76-78
    pub fn from_covariance(cov: Arr, xc: Arr) -> Self {
        Self::new_with_matrix(1.0, cov, xc)
    }
    pub fn kappa(&self) -> f64 {
        let mut v = 0.0;
        check(unsafe { sys::ellcuda_kappa(self.h, &mut v) });
        v
    }
    pub fn mq(&self) -> Arr {
        let mut m = vec![0.0; self.n * self.n];
        check(unsafe { sys::ellcuda_get_mq(self.h, m.as_mut_ptr()) });
        Arr::with_data(m, self.n, self.n)
    }
    fn update(&mut self, which: i32, grad: &Arr, beta0: f64, has_beta1: i32, beta1: f64) -> CutStatus {
        assert_eq!(grad.len(), self.n);
        let mut st: i32 = 3;
        check(unsafe { sys::ellcuda_update(self.h, which, grad.data().as_ptr(), beta0, has_beta1, beta1, &mut st) });
        match st {
            0 => CutStatus::Success,
            1 => CutStatus::NoSoln,
            2 => CutStatus::NoEffect,
            _ => CutStatus::Unknown,
        }
    }
}

impl<const KIND: i32> Clone for EllCudaT<KIND> {
    fn clone(&self) -> Self {
        let mut h: sys::ell_handle = std::ptr::null_mut();
        check(unsafe { sys::ellcuda_clone(self.h, &mut h) });
        EllCudaT { h, n: self.n }
    }
}

impl<const KIND: i32> Drop for EllCudaT<KIND> {
    fn drop(&mut self) {
        unsafe { sys::ellcuda_destroy(self.h) }
    }
}

/// src/ell.rs:140-180
impl<const KIND: i32> SearchSpace for EllCudaT<KIND> {
    type ArrayType = Arr;
    fn xc(&self) -> Arr {
        let mut x = vec![0.0; self.n];
        check(unsafe { sys::ellcuda_xc(self.h, x.as_mut_ptr()) });
        Arr::from(x)
    }
    fn tsq(&self) -> f64 {
        let mut v = 0.0;
        check(unsafe { sys::ellcuda_tsq(self.h, &mut v) });
        v
    }
    fn update_bias_cut<T>(&mut self, cut: &(Arr, T)) -> CutStatus
    where
        T: UpdateByCutChoice<Self, ArrayType = Arr>,
    {
        cut.1.update_bias_cut_by(self, &cut.0)
    }
    fn update_central_cut<T>(&mut self, cut: &(Arr, T)) -> CutStatus
    where
        T: UpdateByCutChoice<Self, ArrayType = Arr>,
    {
        cut.1.update_central_cut_by(self, &cut.0)
    }
    fn update_q<T>(&mut self, cut: &(Arr, T)) -> CutStatus
    where
        T: UpdateByCutChoice<Self, ArrayType = Arr>,
    {
        cut.1.update_q_by(self, &cut.0)
    }
    fn set_xc(&mut self, x: Arr) {
        assert_eq!(x.len(), self.n);
        check(unsafe { sys::ellcuda_set_xc(self.h, x.data().as_ptr()) });
    }
}

/// src/ell.rs:185-196 + :212-229.  The orphan rule allows this impl: the local type is the trait's parameter.
impl<const KIND: i32> UpdateByCutChoice<EllCudaT<KIND>> for SingleCut {
    type ArrayType = Arr;
    fn update_bias_cut_by(&self, e: &mut EllCudaT<KIND>, grad: &Arr) -> CutStatus {
        e.update(sys::ELL_UPDATE_BIAS, grad, self.0, -1, 0.0)
    }
    fn update_central_cut_by(&self, e: &mut EllCudaT<KIND>, grad: &Arr) -> CutStatus {
        e.update(sys::ELL_UPDATE_CENTRAL, grad, self.0, -1, 0.0)
    }
    fn update_q_by(&self, e: &mut EllCudaT<KIND>, grad: &Arr) -> CutStatus {
        e.update(sys::ELL_UPDATE_Q, grad, self.0, -1, 0.0)
    }
}

/// src/ell.rs:198-210 + :212-229
impl<const KIND: i32> UpdateByCutChoice<EllCudaT<KIND>> for ParallelCut {
    type ArrayType = Arr;
    fn update_bias_cut_by(&self, e: &mut EllCudaT<KIND>, grad: &Arr) -> CutStatus {
        e.update(sys::ELL_UPDATE_BIAS, grad, self.0, self.1.is_some() as i32, self.1.unwrap_or(0.0))
    }
    fn update_central_cut_by(&self, e: &mut EllCudaT<KIND>, grad: &Arr) -> CutStatus {
        e.update(sys::ELL_UPDATE_CENTRAL, grad, self.0, self.1.is_some() as i32, self.1.unwrap_or(0.0))
    }
    fn update_q_by(&self, e: &mut EllCudaT<KIND>, grad: &Arr) -> CutStatus {
        e.update(sys::ELL_UPDATE_Q, grad, self.0, self.1.is_some() as i32, self.1.unwrap_or(0.0))
    }
}
