//! Trait-compatible device backend for `rusty-compression` (SOURCE ONLY — never compiled in the
//! authoring environment, which has no Rust toolchain; see ../README.md).
//!
//! `DeviceMatrix<A>` is the device-resident operator.  It implements the reference's plugin API
//! (`MatVec`, `MatMat`, `ConjMatVec`, `ConjMatMat`; reference src/types.rs:40-101) and offers the
//! sampling / QR / SVD entry points with the reference's signatures, forwarding to the C ABI of
//! librc_b200.so (include/rc_api.h).  Because the reference blanket-implements `MatMat` /
//! `ConjMatMat` for every `MatVec` / `ConjMatVec` (src/types.rs:145-146, quirk Q2) and its samplers
//! for every `Op: MatMat`, the resident fast paths are inherent methods with the same names: a
//! caller switches by constructing a `DeviceMatrix` from its `Array2`.

use ndarray::{Array1, Array2, ArrayBase, ArrayView1, ArrayView2, Axis, Data, Ix2};
use rand::Rng;
use rc_b200_sys as sys;
use rusty_compression::types::{c32, c64, ConjMatVec, MatVec, Result, RustyCompressionError, Scalar};
use rusty_compression::{QR, SVD};
use std::marker::PhantomData;
use std::os::raw::c_int;
use std::ptr;
use std::sync::Arc;

pub mod decomp;
pub use decomp::{compute_svd, pivoted_lq, pivoted_qr, DeviceColumnID, DeviceLQ, DeviceQR, DeviceRowID, DeviceSVD, DeviceTwoSidedID};

/// Scalars the library is built for (src/types.rs:9).
pub trait RcScalar: Scalar { const DTYPE: c_int; }
impl RcScalar for f32 { const DTYPE: c_int = sys::RC_F32; }
impl RcScalar for f64 { const DTYPE: c_int = sys::RC_F64; }
impl RcScalar for c32 { const DTYPE: c_int = sys::RC_C32; }
impl RcScalar for c64 { const DTYPE: c_int = sys::RC_C64; }

pub struct Context { pub(crate) raw: *mut sys::rc_ctx }
unsafe impl Send for Context {}
impl Context {
    pub fn new(device: i32) -> Arc<Self> {
        let mut raw = ptr::null_mut();
        let st = unsafe { sys::rc_ctx_create(device, &mut raw) };
        assert_eq!(st, sys::RC_OK, "rc_ctx_create failed (no B200 visible?)");
        Arc::new(Context { raw })
    }
    pub(crate) fn check(&self, st: c_int) -> Result<()> {
        match st {
            sys::RC_OK => Ok(()),
            sys::RC_COMPRESSION_ERROR => Err(RustyCompressionError::CompressionError),
            sys::RC_LAYOUT_ERROR => Err(RustyCompressionError::LayoutError),
            sys::RC_PIVOTED_QR_ERROR => Err(RustyCompressionError::PivotedQRError),
            // where the crate asserts, the ABI returns INVALID_ARGUMENT: keep the panic semantics
            sys::RC_INVALID_ARGUMENT => panic!("{}", self.last_error()),
            _ => panic!("device error {}: {}", st, self.last_error()),
        }
    }
    fn last_error(&self) -> String {
        unsafe { std::ffi::CStr::from_ptr(sys::rc_last_error_string(self.raw)).to_string_lossy().into_owned() }
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { sys::rc_ctx_destroy(self.raw); } } }

pub struct DeviceMatrix<A: RcScalar> { pub(crate) ctx: Arc<Context>, pub(crate) h: *mut sys::rc_matrix, _a: PhantomData<A> }
impl<A: RcScalar> Drop for DeviceMatrix<A> { fn drop(&mut self) { unsafe { sys::rc_matrix_free(self.h); } } }

/// An upload in flight (`DeviceMatrix::from_array_async`); borrows the host array until it is consumed.
pub struct PendingMatrix<'a, A: RcScalar> { m: DeviceMatrix<A>, _src: PhantomData<&'a A> }
impl<'a, A: RcScalar> PendingMatrix<'a, A> {
    /// Orders the context stream behind the copy and waits for it on the host (the borrow of the source ends here).
    pub fn ready(self) -> Result<DeviceMatrix<A>> {
        self.m.ctx.check(unsafe { sys::rc_matrix_await(self.m.ctx.raw, self.m.h, 1) })?;
        Ok(self.m)
    }
}

impl<A: RcScalar> DeviceMatrix<A> {
    /// Upload any ndarray view; the ABI takes element strides (like `mat.assign(&arr)`, src/pivoted_qr.rs:29).
    pub fn from_array<S: Data<Elem = A>>(ctx: &Arc<Context>, a: &ArrayBase<S, Ix2>) -> Result<Self> {
        let (rs, cs) = (a.strides()[0] as i64, a.strides()[1] as i64);
        let mut h = ptr::null_mut();
        ctx.check(unsafe { sys::rc_matrix_from_host(ctx.raw, A::DTYPE, a.as_ptr() as *const _, a.nrows() as i64,
                                                    a.ncols() as i64, rs, cs, &mut h) })?;
        Ok(DeviceMatrix { ctx: ctx.clone(), h, _a: PhantomData })
    }
    /// Pipelined upload of a row-major (standard layout) array that the caller keeps alive -- and pinned, for a truly
    /// asynchronous copy -- until `ready`: the transfer runs on the context's copy stream under the kernels of the
    /// current step.  The pending handle can only be used through `ready`, which orders the context stream behind it.
    pub fn from_array_async<'a>(ctx: &Arc<Context>, a: &'a ArrayView2<'a, A>) -> Result<PendingMatrix<'a, A>> {
        assert!(a.strides()[1] == 1 && a.strides()[0] >= a.ncols() as isize, "from_array_async takes a row-major view");
        let mut h = ptr::null_mut();
        ctx.check(unsafe { sys::rc_matrix_from_host_async(ctx.raw, A::DTYPE, a.as_ptr() as *const _, a.nrows() as i64,
                                                          a.ncols() as i64, a.strides()[0] as i64, &mut h) })?;
        Ok(PendingMatrix { m: DeviceMatrix { ctx: ctx.clone(), h, _a: PhantomData }, _src: PhantomData })
    }
    pub(crate) fn from_raw(ctx: &Arc<Context>, h: *mut sys::rc_matrix) -> Self { DeviceMatrix { ctx: ctx.clone(), h, _a: PhantomData } }
    pub(crate) fn download(ctx: &Arc<Context>, h: *const sys::rc_matrix) -> Result<Array2<A>> {
        let (r, c) = unsafe { (sys::rc_matrix_rows(h) as usize, sys::rc_matrix_cols(h) as usize) };
        let mut out = Array2::<A>::zeros((r, c));
        ctx.check(unsafe { sys::rc_matrix_to_host(ctx.raw, h, out.as_mut_ptr() as *mut _) })?;
        Ok(out)
    }
    pub fn to_array(&self) -> Result<Array2<A>> { Self::download(&self.ctx, self.h) }

    /// MatMat::matmat (src/types.rs:58-71): ONE GEMM on the tensor pipe instead of a GEMV per column.
    pub fn matmat_device(&self, x: &DeviceMatrix<A>) -> Result<DeviceMatrix<A>> {
        let mut y = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_matmat(self.ctx.raw, self.h, x.h, &mut y) })?;
        Ok(Self::from_raw(&self.ctx, y))
    }
    /// ConjMatMat::conj_matmat (src/types.rs:88-101).
    pub fn conj_matmat_device(&self, x: &DeviceMatrix<A>) -> Result<DeviceMatrix<A>> {
        let mut z = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_conj_matmat(self.ctx.raw, self.h, x.h, &mut z) })?;
        Ok(Self::from_raw(&self.ctx, z))
    }

    /// SampleRange::sample_range_by_rank (src/random_sampling.rs:103-118); Omega from Philox seeded by `rng`.
    pub fn sample_range_by_rank<R: Rng>(&self, k: usize, p: usize, rng: &mut R) -> Result<Array2<A>> {
        let mut q = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_sample_range_by_rank(self.ctx.raw, self.h, k as i64, p as i64, ptr::null(),
                                                             rng.next_u64(), &mut q) })?;
        Self::from_raw(&self.ctx, q).to_array()
    }
    /// SampleRangePowerIteration::sample_range_power_iteration (src/random_sampling.rs:131-160, quirk Q1 included).
    pub fn sample_range_power_iteration<R: Rng>(&self, k: usize, p: usize, it_count: usize, rng: &mut R) -> Result<Array2<A>> {
        let mut q = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_sample_range_power_iteration(self.ctx.raw, self.h, k as i64, p as i64,
                                                                     it_count as i64, ptr::null(), rng.next_u64(), &mut q) })?;
        Self::from_raw(&self.ctx, q).to_array()
    }
    /// AdaptiveSampling::sample_range_adaptive (src/random_sampling.rs:223-274).
    pub fn sample_range_adaptive<R: Rng>(&self, rel_tol: f64, sample_size: usize, rng: &mut R)
        -> Result<(Array2<A>, Vec<(usize, f64)>)> {
        let (mut q, mut n) = (ptr::null_mut(), 0usize);
        let (mut ranks, mut res) = (vec![0u64; 4096], vec![0f64; 4096]);
        self.ctx.check(unsafe { sys::rc_sample_range_adaptive(self.ctx.raw, self.h, rel_tol, sample_size as i64, ptr::null(),
            rng.next_u64(), 0, &mut q, ranks.as_mut_ptr(), res.as_mut_ptr(), ranks.len(), &mut n) })?;
        let hist = (0..n.min(4096)).map(|i| (ranks[i] as usize, res[i])).collect();
        Ok((Self::from_raw(&self.ctx, q).to_array()?, hist))
    }
    /// QRTraits::compute_from_range_estimate (src/qr.rs:311-323) with the operator resident on device.
    pub fn qr_from_range_estimate(&self, range: ArrayView2<A>) -> Result<QR<A>> {
        let rd = DeviceMatrix::from_array(&self.ctx, &range)?;
        let mut h = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_qr_compute_from_range_estimate(self.ctx.raw, rd.h, self.h, &mut h) })?;
        let n = unsafe { sys::rc_qr_ncols(h) as usize };
        let mut ind = vec![0u64; n];
        self.ctx.check(unsafe { sys::rc_qr_get_ind(h, ind.as_mut_ptr(), n) })?;
        let out = QR { q: Self::download(&self.ctx, unsafe { sys::rc_qr_get_q(h) })?,
                       r: Self::download(&self.ctx, unsafe { sys::rc_qr_get_r(h) })?,
                       ind: Array1::from(ind.into_iter().map(|i| i as usize).collect::<Vec<_>>()) };
        unsafe { sys::rc_qr_free(h); }
        Ok(out)
    }
    /// SVDTraits::compute_from_range_estimate (src/svd.rs:171-183).
    pub fn svd_from_range_estimate(&self, range: ArrayView2<A>) -> Result<SVD<A>> where A::Real: From<f64> {
        let rd = DeviceMatrix::from_array(&self.ctx, &range)?;
        let mut h = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_svd_compute_from_range_estimate(self.ctx.raw, rd.h, self.h, &mut h) })?;
        let k = unsafe { sys::rc_svd_rank(h) as usize };
        let mut s = vec![0f64; k];
        self.ctx.check(unsafe { sys::rc_svd_get_s(h, s.as_mut_ptr(), k) })?;
        let out = SVD { u: Self::download(&self.ctx, unsafe { sys::rc_svd_get_u(h) })?,
                        s: Array1::from(s.into_iter().map(A::Real::from).collect::<Vec<_>>()),
                        vt: Self::download(&self.ctx, unsafe { sys::rc_svd_get_vt(h) })? };
        unsafe { sys::rc_svd_free(h); }
        Ok(out)
    }
}

// The plugin API itself, so a DeviceMatrix can be handed to any code written against the reference's traits.
impl<A: RcScalar> MatVec for DeviceMatrix<A> {
    type A = A;
    fn nrows(&self) -> usize { unsafe { sys::rc_matrix_rows(self.h) as usize } }
    fn ncols(&self) -> usize { unsafe { sys::rc_matrix_cols(self.h) as usize } }
    fn matvec(&self, x: ArrayView1<A>) -> Array1<A> {
        let xd = DeviceMatrix::from_array(&self.ctx, &x.insert_axis(Axis(1))).expect("upload");
        self.matmat_device(&xd).and_then(|y| y.to_array()).expect("rc_matmat").column(0).to_owned()
    }
}
impl<A: RcScalar> ConjMatVec for DeviceMatrix<A> {
    fn conj_matvec(&self, x: ArrayView1<A>) -> Array1<A> {
        let xd = DeviceMatrix::from_array(&self.ctx, &x.insert_axis(Axis(1))).expect("upload");
        self.conj_matmat_device(&xd).and_then(|y| y.to_array()).expect("rc_conj_matmat").column(0).to_owned()
    }
}

// ------------------------------------------------------------------------------------------------
// Matrix-free operators: the crate's plugin API proper (anything implementing MatVec / ConjMatVec,
// src/types.rs:40-51, 77-81) for operators that are never materialised.  A user type that can launch
// its products on a CUDA stream implements `DeviceOperator`; `OperatorHandle::new` turns it into an
// `rc_matrix*` (rc_operator_create) that every operator-taking entry point of the C ABI accepts.

/// Products of a user operator on DEVICE buffers (row-major, leading dimensions in elements),
/// enqueued on `stream` (the cudaStream_t the library works on).  Return 0 on success.
pub trait DeviceOperator<A: RcScalar> {
    fn nrows(&self) -> usize;
    fn ncols(&self) -> usize;
    /// y (nrows x ncols_x) = A x,  x is ncols x ncols_x
    unsafe fn matmat(&self, x: *const A, ldx: i64, ncols_x: i64, y: *mut A, ldy: i64, stream: *mut std::os::raw::c_void) -> c_int;
    /// z (ncols x ncols_x) = A^H x,  x is nrows x ncols_x
    unsafe fn conj_matmat(&self, x: *const A, ldx: i64, ncols_x: i64, z: *mut A, ldz: i64, stream: *mut std::os::raw::c_void) -> c_int;
}

pub struct OperatorHandle<A: RcScalar, Op: DeviceOperator<A>> {
    ctx: Arc<Context>,
    h: *mut sys::rc_matrix,
    _op: Box<Op>,                       // the callbacks borrow it through `user`
    _a: PhantomData<A>,
}
impl<A: RcScalar, Op: DeviceOperator<A>> Drop for OperatorHandle<A, Op> {
    fn drop(&mut self) { unsafe { sys::rc_matrix_free(self.h); } }
}

unsafe extern "C" fn tramp_matmat<A: RcScalar, Op: DeviceOperator<A>>(user: *mut std::os::raw::c_void, x: *const std::os::raw::c_void,
        ldx: i64, ncols: i64, y: *mut std::os::raw::c_void, ldy: i64, stream: *mut std::os::raw::c_void) -> c_int {
    (&*(user as *const Op)).matmat(x as *const A, ldx, ncols, y as *mut A, ldy, stream)
}
unsafe extern "C" fn tramp_conj<A: RcScalar, Op: DeviceOperator<A>>(user: *mut std::os::raw::c_void, x: *const std::os::raw::c_void,
        ldx: i64, ncols: i64, z: *mut std::os::raw::c_void, ldz: i64, stream: *mut std::os::raw::c_void) -> c_int {
    (&*(user as *const Op)).conj_matmat(x as *const A, ldx, ncols, z as *mut A, ldz, stream)
}

impl<A: RcScalar, Op: DeviceOperator<A>> OperatorHandle<A, Op> {
    pub fn new(ctx: &Arc<Context>, op: Op) -> Result<Self> {
        let op = Box::new(op);
        let mut h = ptr::null_mut();
        ctx.check(unsafe {
            sys::rc_operator_create(ctx.raw, A::DTYPE, op.nrows() as i64, op.ncols() as i64,
                                    Some(tramp_matmat::<A, Op>), Some(tramp_conj::<A, Op>),
                                    &*op as *const Op as *mut std::os::raw::c_void, &mut h)
        })?;
        Ok(Self { ctx: ctx.clone(), h, _op: op, _a: PhantomData })
    }

    /// SampleRange::sample_range_by_rank (src/random_sampling.rs:103-118) on the matrix-free operator.
    pub fn sample_range_by_rank<R: Rng>(&self, k: usize, p: usize, rng: &mut R) -> Result<DeviceMatrix<A>> {
        let mut q = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_sample_range_by_rank(self.ctx.raw, self.h, k as i64, p as i64, ptr::null(), rng.next_u64(), &mut q) })?;
        Ok(DeviceMatrix { ctx: self.ctx.clone(), h: q, _a: PhantomData })
    }
    /// Raw handle for the other operator-taking entry points (rc_sample_range_power_iteration,
    /// rc_sample_range_adaptive, rc_qr_compute_from_range_estimate, rc_svd_compute_from_range_estimate).
    pub fn raw(&self) -> *const sys::rc_matrix { self.h }
}
