//! Device-resident decompositions and the two `pub(crate)` seams of the reference (SOURCE ONLY, see ../README.md).
//!
//! * `pivoted_qr` / `pivoted_lq` / `compute_svd` take and return host `ndarray`s with the signatures of the
//!   reference's internal seams (`PivotedQR::pivoted_qr`, src/pivoted_qr.rs:11-19; `ComputeSVD::compute_svd`,
//!   src/compute_svd.rs:8-12): swapping those two trait impls for these functions moves every LAPACK call of the
//!   crate (`?geqp3`, `?orgqr/?ungqr`, `?gesdd`) onto the B200 without touching the rest of its code.
//! * `DeviceQR`, `DeviceLQ`, `DeviceSVD`, `DeviceColumnID`, `DeviceRowID`, `DeviceTwoSidedID` keep the factors in
//!   HBM and offer the methods of `QRTraits`, `LQTraits`, `SVDTraits`, `ColumnIDTraits`, `RowIDTraits`,
//!   `TwoSidedIDTraits` and `Apply` (src/qr.rs:54-237, src/svd.rs:23-122, src/col_interp_decomp.rs:44-86,
//!   src/row_interp_decomp.rs:46-89, src/two_sided_interp_decomp.rs:43-96); `to_host()` converts to the
//!   reference's own structs (`QR`, `LQ`, `SVD`, `ColumnID`, `RowID`, `TwoSidedID`).

use crate::{Context, DeviceMatrix, RcScalar};
use ndarray::{Array1, Array2, ArrayView2};
use rc_b200_sys as sys;
use rusty_compression::types::Result;
use rusty_compression::{ColumnID, ColumnIDTraits, CompressionType, RowID, RowIDTraits, TwoSidedID, TwoSidedIDTraits, LQ, QR, SVD};
use std::marker::PhantomData;
use std::ptr;
use std::sync::Arc;

fn usize_vec(v: Vec<u64>) -> Array1<usize> { Array1::from(v.into_iter().map(|i| i as usize).collect::<Vec<_>>()) }

macro_rules! device_handle {
    ($name:ident, $raw:ident, $free:ident) => {
        pub struct $name<A: RcScalar> { pub(crate) ctx: Arc<Context>, pub(crate) h: *mut sys::$raw, _a: PhantomData<A> }
        impl<A: RcScalar> Drop for $name<A> { fn drop(&mut self) { unsafe { sys::$free(self.h); } } }
        impl<A: RcScalar> $name<A> {
            pub(crate) fn from_raw(ctx: &Arc<Context>, h: *mut sys::$raw) -> Self { $name { ctx: ctx.clone(), h, _a: PhantomData } }
        }
    };
}
device_handle!(DeviceQR, rc_qr, rc_qr_free);
device_handle!(DeviceLQ, rc_lq, rc_lq_free);
device_handle!(DeviceSVD, rc_svd, rc_svd_free);
device_handle!(DeviceColumnID, rc_column_id, rc_column_id_free);
device_handle!(DeviceRowID, rc_row_id, rc_row_id_free);
device_handle!(DeviceTwoSidedID, rc_two_sided_id, rc_two_sided_id_free);

/// A new owned device matrix produced by a `rc_*_to_mat` / `rc_*_apply` style call.
macro_rules! new_matrix {
    ($self:ident, $call:ident $(, $arg:expr)*) => {{
        let mut out = ptr::null_mut();
        $self.ctx.check(unsafe { sys::$call($self.ctx.raw, $self.h $(, $arg)*, &mut out) })?;
        Ok(DeviceMatrix::from_raw(&$self.ctx, out))
    }};
}

impl<A: RcScalar> DeviceQR<A> {
    /// `QR { q, r, ind }` assembled from parts (the reference's struct has pub fields, src/qr.rs:31-40).
    pub fn from_parts(q: &DeviceMatrix<A>, r: &DeviceMatrix<A>, ind: &[usize]) -> Result<Self> {
        let iv: Vec<u64> = ind.iter().map(|&i| i as u64).collect();
        let mut h = ptr::null_mut();
        q.ctx.check(unsafe { sys::rc_qr_new(q.ctx.raw, q.h, r.h, iv.as_ptr(), iv.len(), &mut h) })?;
        Ok(Self::from_raw(&q.ctx, h))
    }
    /// QRTraits::compute_from (src/qr.rs:214, 251-253)
    pub fn compute_from(arr: &DeviceMatrix<A>) -> Result<Self> {
        let mut h = ptr::null_mut();
        arr.ctx.check(unsafe { sys::rc_qr_compute_from(arr.ctx.raw, arr.h, &mut h) })?;
        Ok(Self::from_raw(&arr.ctx, h))
    }
    /// QRTraits::compute_from_range_estimate (src/qr.rs:221-224, 311-323); `op` dense or matrix-free.
    pub fn compute_from_range_estimate(range: &DeviceMatrix<A>, op: &DeviceMatrix<A>) -> Result<Self> {
        let mut h = ptr::null_mut();
        op.ctx.check(unsafe { sys::rc_qr_compute_from_range_estimate(op.ctx.raw, range.h, op.h, &mut h) })?;
        Ok(Self::from_raw(&op.ctx, h))
    }
    pub fn nrows(&self) -> usize { unsafe { sys::rc_qr_nrows(self.h) as usize } }
    pub fn ncols(&self) -> usize { unsafe { sys::rc_qr_ncols(self.h) as usize } }
    pub fn rank(&self) -> usize { unsafe { sys::rc_qr_rank(self.h) as usize } }
    pub fn get_ind(&self) -> Result<Array1<usize>> {
        let mut v = vec![0u64; self.ncols()];
        self.ctx.check(unsafe { sys::rc_qr_get_ind(self.h, v.as_mut_ptr(), v.len()) })?;
        Ok(usize_vec(v))
    }
    pub fn to_mat(&self) -> Result<DeviceMatrix<A>> { new_matrix!(self, rc_qr_to_mat) }
    /// QRTraits::compress (src/qr.rs:169-208); ADAPTIVE errors with CompressionError when no |r_ii / r_00| < tol.
    pub fn compress(&self, compression_type: CompressionType) -> Result<Self> {
        let mut out = ptr::null_mut();
        let st = match compression_type {
            CompressionType::RANK(k) => unsafe { sys::rc_qr_compress_rank(self.ctx.raw, self.h, k as i64, &mut out) },
            CompressionType::ADAPTIVE(tol) => unsafe { sys::rc_qr_compress_tolerance(self.ctx.raw, self.h, tol, &mut out) },
        };
        self.ctx.check(st)?;
        Ok(Self::from_raw(&self.ctx, out))
    }
    /// QRTraits::column_id (src/qr.rs:270-309)
    pub fn column_id(&self) -> Result<DeviceColumnID<A>> {
        let mut out = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_qr_column_id(self.ctx.raw, self.h, &mut out) })?;
        Ok(DeviceColumnID::from_raw(&self.ctx, out))
    }
    pub fn to_host(&self) -> Result<QR<A>> {
        Ok(QR { q: DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_qr_get_q(self.h) })?,
                r: DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_qr_get_r(self.h) })?,
                ind: self.get_ind()? })
    }
}

impl<A: RcScalar> DeviceLQ<A> {
    /// `LQ { l, q, ind }` assembled from parts (pub fields, src/qr.rs:42-51).
    pub fn from_parts(l: &DeviceMatrix<A>, q: &DeviceMatrix<A>, ind: &[usize]) -> Result<Self> {
        let iv: Vec<u64> = ind.iter().map(|&i| i as u64).collect();
        let mut h = ptr::null_mut();
        l.ctx.check(unsafe { sys::rc_lq_new(l.ctx.raw, l.h, q.h, iv.as_ptr(), iv.len(), &mut h) })?;
        Ok(Self::from_raw(&l.ctx, h))
    }
    /// LQTraits::compute_from (src/qr.rs:135, 354-362)
    pub fn compute_from(arr: &DeviceMatrix<A>) -> Result<Self> {
        let mut h = ptr::null_mut();
        arr.ctx.check(unsafe { sys::rc_lq_compute_from(arr.ctx.raw, arr.h, &mut h) })?;
        Ok(Self::from_raw(&arr.ctx, h))
    }
    pub fn nrows(&self) -> usize { unsafe { sys::rc_lq_nrows(self.h) as usize } }
    pub fn ncols(&self) -> usize { unsafe { sys::rc_lq_ncols(self.h) as usize } }
    pub fn rank(&self) -> usize { unsafe { sys::rc_lq_rank(self.h) as usize } }
    pub fn get_ind(&self) -> Result<Array1<usize>> {
        let mut v = vec![0u64; self.nrows()];
        self.ctx.check(unsafe { sys::rc_lq_get_ind(self.h, v.as_mut_ptr(), v.len()) })?;
        Ok(usize_vec(v))
    }
    pub fn to_mat(&self) -> Result<DeviceMatrix<A>> { new_matrix!(self, rc_lq_to_mat) }
    /// LQTraits::compress (src/qr.rs:80-119)
    pub fn compress(&self, compression_type: CompressionType) -> Result<Self> {
        let mut out = ptr::null_mut();
        let st = match compression_type {
            CompressionType::RANK(k) => unsafe { sys::rc_lq_compress_rank(self.ctx.raw, self.h, k as i64, &mut out) },
            CompressionType::ADAPTIVE(tol) => unsafe { sys::rc_lq_compress_tolerance(self.ctx.raw, self.h, tol, &mut out) },
        };
        self.ctx.check(st)?;
        Ok(Self::from_raw(&self.ctx, out))
    }
    /// LQTraits::row_id (src/qr.rs:363-403)
    pub fn row_id(&self) -> Result<DeviceRowID<A>> {
        let mut out = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_lq_row_id(self.ctx.raw, self.h, &mut out) })?;
        Ok(DeviceRowID::from_raw(&self.ctx, out))
    }
    pub fn to_host(&self) -> Result<LQ<A>> {
        Ok(LQ { l: DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_lq_get_l(self.h) })?,
                q: DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_lq_get_q(self.h) })?,
                ind: self.get_ind()? })
    }
}

impl<A: RcScalar> DeviceSVD<A> where A::Real: From<f64> {
    /// SVDTraits::compute_from (src/svd.rs:103, 165-169)
    pub fn compute_from(arr: &DeviceMatrix<A>) -> Result<Self> {
        let mut h = ptr::null_mut();
        arr.ctx.check(unsafe { sys::rc_svd_compute_from(arr.ctx.raw, arr.h, &mut h) })?;
        Ok(Self::from_raw(&arr.ctx, h))
    }
    /// SVDTraits::compute_from_range_estimate (src/svd.rs:110-113, 171-183)
    pub fn compute_from_range_estimate(range: &DeviceMatrix<A>, op: &DeviceMatrix<A>) -> Result<Self> {
        let mut h = ptr::null_mut();
        op.ctx.check(unsafe { sys::rc_svd_compute_from_range_estimate(op.ctx.raw, range.h, op.h, &mut h) })?;
        Ok(Self::from_raw(&op.ctx, h))
    }
    pub fn rank(&self) -> usize { unsafe { sys::rc_svd_rank(self.h) as usize } }
    pub fn get_s(&self) -> Result<Array1<A::Real>> {
        let mut s = vec![0f64; self.rank()];
        self.ctx.check(unsafe { sys::rc_svd_get_s(self.h, s.as_mut_ptr(), s.len()) })?;
        Ok(Array1::from(s.into_iter().map(A::Real::from).collect::<Vec<_>>()))
    }
    pub fn to_mat(&self) -> Result<DeviceMatrix<A>> { new_matrix!(self, rc_svd_to_mat) }
    /// SVDTraits::to_qr (src/svd.rs:57, 150-163)
    pub fn to_qr(&self) -> Result<DeviceQR<A>> {
        let mut out = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_svd_to_qr(self.ctx.raw, self.h, &mut out) })?;
        Ok(DeviceQR::from_raw(&self.ctx, out))
    }
    /// SVDTraits::compress (src/svd.rs:60-101)
    pub fn compress(&self, compression_type: CompressionType) -> Result<Self> {
        let mut out = ptr::null_mut();
        let st = match compression_type {
            CompressionType::RANK(k) => unsafe { sys::rc_svd_compress_rank(self.ctx.raw, self.h, k as i64, &mut out) },
            CompressionType::ADAPTIVE(tol) => unsafe { sys::rc_svd_compress_tolerance(self.ctx.raw, self.h, tol, &mut out) },
        };
        self.ctx.check(st)?;
        Ok(Self::from_raw(&self.ctx, out))
    }
    pub fn to_host(&self) -> Result<SVD<A>> {
        Ok(SVD { u: DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_svd_get_u(self.h) })?,
                 s: self.get_s()?,
                 vt: DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_svd_get_vt(self.h) })? })
    }
}

impl<A: RcScalar> DeviceColumnID<A> where ColumnID<A>: ColumnIDTraits<A = A> {
    pub fn get_col_ind(&self) -> Result<Array1<usize>> {
        let mut v = vec![0u64; unsafe { sys::rc_column_id_col_ind_len(self.h) }];
        self.ctx.check(unsafe { sys::rc_column_id_get_col_ind(self.h, v.as_mut_ptr(), v.len()) })?;
        Ok(usize_vec(v))
    }
    pub fn to_mat(&self) -> Result<DeviceMatrix<A>> { new_matrix!(self, rc_column_id_to_mat) }
    /// Apply::dot (src/col_interp_decomp.rs:134-154), all right-hand sides at once
    pub fn dot(&self, rhs: &DeviceMatrix<A>) -> Result<DeviceMatrix<A>> { new_matrix!(self, rc_column_id_apply, rhs.h) }
    /// ColumnIDTraits::two_sided_id (src/col_interp_decomp.rs:85, 116-125)
    pub fn two_sided_id(&self) -> Result<DeviceTwoSidedID<A>> {
        let mut out = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_column_id_two_sided_id(self.ctx.raw, self.h, &mut out) })?;
        Ok(DeviceTwoSidedID::from_raw(&self.ctx, out))
    }
    pub fn to_host(&self) -> Result<ColumnID<A>> {
        Ok(ColumnID::<A>::new(DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_column_id_get_c(self.h) })?,
                              DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_column_id_get_z(self.h) })?,
                              self.get_col_ind()?))
    }
}

impl<A: RcScalar> DeviceRowID<A> where RowID<A>: RowIDTraits<A = A> {
    pub fn get_row_ind(&self) -> Result<Array1<usize>> {
        let mut v = vec![0u64; unsafe { sys::rc_row_id_row_ind_len(self.h) }];
        self.ctx.check(unsafe { sys::rc_row_id_get_row_ind(self.h, v.as_mut_ptr(), v.len()) })?;
        Ok(usize_vec(v))
    }
    pub fn to_mat(&self) -> Result<DeviceMatrix<A>> { new_matrix!(self, rc_row_id_to_mat) }
    /// Apply::dot (src/row_interp_decomp.rs:134-154)
    pub fn dot(&self, rhs: &DeviceMatrix<A>) -> Result<DeviceMatrix<A>> { new_matrix!(self, rc_row_id_apply, rhs.h) }
    /// RowIDTraits::two_sided_id (src/row_interp_decomp.rs:87, 120-130)
    pub fn two_sided_id(&self) -> Result<DeviceTwoSidedID<A>> {
        let mut out = ptr::null_mut();
        self.ctx.check(unsafe { sys::rc_row_id_two_sided_id(self.ctx.raw, self.h, &mut out) })?;
        Ok(DeviceTwoSidedID::from_raw(&self.ctx, out))
    }
    pub fn to_host(&self) -> Result<RowID<A>> {
        Ok(RowID::<A>::new(DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_row_id_get_x(self.h) })?,
                           DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_row_id_get_r(self.h) })?,
                           self.get_row_ind()?))
    }
}

impl<A: RcScalar> DeviceTwoSidedID<A> where TwoSidedID<A>: TwoSidedIDTraits<A = A> {
    pub fn get_col_ind(&self) -> Result<Array1<usize>> {
        let mut v = vec![0u64; unsafe { sys::rc_two_sided_id_col_ind_len(self.h) }];
        self.ctx.check(unsafe { sys::rc_two_sided_id_get_col_ind(self.h, v.as_mut_ptr(), v.len()) })?;
        Ok(usize_vec(v))
    }
    pub fn get_row_ind(&self) -> Result<Array1<usize>> {
        let mut v = vec![0u64; unsafe { sys::rc_two_sided_id_row_ind_len(self.h) }];
        self.ctx.check(unsafe { sys::rc_two_sided_id_get_row_ind(self.h, v.as_mut_ptr(), v.len()) })?;
        Ok(usize_vec(v))
    }
    pub fn to_mat(&self) -> Result<DeviceMatrix<A>> { new_matrix!(self, rc_two_sided_id_to_mat) }
    /// Apply::dot (src/two_sided_interp_decomp.rs:154-171)
    pub fn dot(&self, rhs: &DeviceMatrix<A>) -> Result<DeviceMatrix<A>> { new_matrix!(self, rc_two_sided_id_apply, rhs.h) }
    /// Argument order of the crate's constructor: (x, r, c, col_ind, row_ind) (src/two_sided_interp_decomp.rs:89-95).
    pub fn to_host(&self) -> Result<TwoSidedID<A>> {
        Ok(TwoSidedID::<A>::new(DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_two_sided_id_get_x(self.h) })?,
                                DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_two_sided_id_get_r(self.h) })?,
                                DeviceMatrix::<A>::download(&self.ctx, unsafe { sys::rc_two_sided_id_get_c(self.h) })?,
                                self.get_col_ind()?, self.get_row_ind()?))
    }
}

// ------------------------------------------------------------------------------------------------
// The two internal seams of the reference, host arrays in and out.

/// `PivotedQR::pivoted_qr` (src/pivoted_qr.rs:11-19, 25-31): `arr[:, ind] = q r`, any layout accepted.
pub fn pivoted_qr<A: RcScalar>(ctx: &Arc<Context>, arr: ArrayView2<A>) -> Result<(Array2<A>, Array2<A>, Array1<usize>)> {
    let qr = DeviceQR::compute_from(&DeviceMatrix::from_array(ctx, &arr)?)?.to_host()?;
    Ok((qr.q, qr.r, qr.ind))
}

/// `PivotedQR::pivoted_lq` (src/pivoted_qr.rs:32-41): `arr[ind, :] = l q`.
pub fn pivoted_lq<A: RcScalar>(ctx: &Arc<Context>, arr: ArrayView2<A>) -> Result<(Array2<A>, Array2<A>, Array1<usize>)> {
    let lq = DeviceLQ::compute_from(&DeviceMatrix::from_array(ctx, &arr)?)?.to_host()?;
    Ok((lq.l, lq.q, lq.ind))
}

/// `ComputeSVD::compute_svd` (src/compute_svd.rs:8-12): thin SVD, singular values descending.
pub fn compute_svd<A: RcScalar>(ctx: &Arc<Context>, arr: ArrayView2<A>) -> Result<(Array2<A>, Array1<A::Real>, Array2<A>)>
where A::Real: From<f64> {
    let svd = DeviceSVD::compute_from(&DeviceMatrix::from_array(ctx, &arr)?)?.to_host()?;
    Ok((svd.u, svd.s, svd.vt))
}
