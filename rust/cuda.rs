// qrusty/src/cuda.rs -- Rust side of the CUDA hot path.  Add `pub mod cuda ;` next to
// `pub mod accel ;` (qrusty/src/lib.rs:29) and the `AccelMode::Cuda` arms shown in
// INTEGRATION.md.  This file binds include/qrusty_cuda.h one to one; it holds no algorithm.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no cargo/rustc.  The same ABI is
// exercised from Python (qrusty_b200/_ffi.py) by the parity tests.

use num_complex::Complex64;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

use crate::accel::UnsafeVectors;
use crate::{PauliSummand, QrustyErr};

/// qr_term: one tuple of `accel::rowwise::make_params` (accel.rs:141-157), `repr(C)`.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct QrTerm {
    pub z: u64,
    pub x: u64,
    pub re: f64,
    pub im: f64,
}

#[repr(C)]
#[derive(Default)]
pub struct QrPlanInfo {
    pub n_qubits: i32,
    pub device: i32,
    pub dim: u64,
    pub n_terms: u64,
    pub n_groups: u64,
    pub nnz: u64,
}

#[repr(C)]
pub struct QrPlan {
    _private: [u8; 0],
}

pub const QR_INDPTR_LOCAL: u32 = 0;
pub const QR_INDPTR_GLOBAL: u32 = 1;

#[link(name = "qrusty_cuda")]
extern "C" {
    fn qr_plan_create(n_qubits: c_int, terms: *const QrTerm, n_terms: usize, device: c_int,
                      flags: u32, out: *mut *mut QrPlan) -> c_int;
    fn qr_plan_destroy(plan: *mut QrPlan) -> c_int;
    fn qr_plan_info(plan: *const QrPlan, info: *mut QrPlanInfo) -> c_int;
    fn qr_build_host(plan: *mut QrPlan, row_lo: u64, row_hi: u64, indptr: *mut u64,
                     indices: *mut u64, data: *mut f64, flags: u32) -> c_int;
    fn qr_apply_host(plan: *mut QrPlan, v: *const f64, y: *mut f64) -> c_int;
    fn qr_build_compact_count(plan: *mut QrPlan, row_lo: u64, row_hi: u64, tol: f64,
                              d_indptr: *mut u64, nnz_out: *mut u64, stream: *mut c_void) -> c_int;
    fn qr_build_compact_fill(plan: *mut QrPlan, row_lo: u64, row_hi: u64, tol: f64, d_indptr: *const u64,
                             d_indices: *mut u64, d_data: *mut f64, stream: *mut c_void) -> c_int;
    fn qr_write_rawio(plan: *mut QrPlan, row_lo: u64, row_hi: u64, path: *const c_char) -> c_int;
    fn qr_precond2_device(n: u64, d_diag: *const f64, d_dx: *const f64, e: *const f64, tol: f64,
                          d_out: *mut f64, stream: *mut c_void) -> c_int;
    fn qr_malloc_device(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    fn qr_free_device(ptr: *mut c_void) -> c_int;
    fn qr_memcpy_h2d(dst: *mut c_void, src: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
    fn qr_memcpy_d2h(dst: *mut c_void, src: *const c_void, bytes: usize, stream: *mut c_void) -> c_int;
    fn qr_device_count(count: *mut c_int) -> c_int;
    fn qr_last_error() -> *const c_char;
    #[allow(dead_code)]
    fn qr_malloc_host(ptr: *mut *mut c_void, bytes: usize) -> c_int;
}

fn last_error() -> QrustyErr {
    // QrustyErr carries a &'static str (lib.rs:34-46); leak the message (errors are rare).
    let msg = unsafe { CStr::from_ptr(qr_last_error()) }.to_string_lossy().into_owned();
    QrustyErr::new(Box::leak(msg.into_boxed_str()))
}

fn check(rc: c_int) -> Result<(), QrustyErr> {
    if rc == 0 { Ok(()) } else { Err(last_error()) }
}

/// The operator, canonicalised and resident on one GPU.
pub struct Plan {
    raw: *mut QrPlan,
    pub info: QrPlanInfo,
}

impl Plan {
    /// `members` as SparsePauliOp holds them (lib.rs:333-348).  The coefficient handed to the
    /// device follows the default `to_matrix`: coeff * (+i)^base_phase * (-i)^#Y
    /// (lib.rs:182-190,211 and accel.rs:147-153); for labels without an i/j prefix this is
    /// exactly `accel::rowwise::make_params`.
    pub fn new(members: &[PauliSummand], device: i32) -> Result<Plan, QrustyErr> {
        let units = [Complex64::new(1.0, 0.0), Complex64::new(0.0, -1.0),
                     Complex64::new(-1.0, 0.0), Complex64::new(0.0, 1.0)];
        let terms: Vec<QrTerm> = members.iter().map(|(p, coeff)| {
            let n_y = (p.phase() + 4 - p.base_phase() % 4) % 4;        // phase() = base_phase + #Y
            let unit = units[(n_y + 4 - p.base_phase() % 4) % 4];
            let c = unit * *coeff;
            QrTerm { z: p.z_indices(), x: p.x_indices(), re: c.re, im: c.im }
        }).collect();
        let n_qubits = members[0].0.num_qubits() as c_int;
        let mut raw: *mut QrPlan = std::ptr::null_mut();
        check(unsafe { qr_plan_create(n_qubits, terms.as_ptr(), terms.len(), device, 0, &mut raw) })?;
        let mut info = QrPlanInfo::default();
        check(unsafe { qr_plan_info(raw, &mut info) })?;
        Ok(Plan { raw, info })
    }

    /// Replaces `accel::rowwise::make_unsafe_vectors_chunked` (accel.rs:267-336): same
    /// UnsafeVectors, rows [row_lo,row_hi), indptr local to the window.
    pub fn build_rows(&self, row_lo: u64, row_hi: u64) -> Result<UnsafeVectors, QrustyErr> {
        let rows = (row_hi - row_lo) as usize;
        let nnz = rows * self.info.n_groups as usize;
        let mut indptr: Vec<u64> = Vec::with_capacity(rows + 1);
        let mut indices: Vec<u64> = Vec::with_capacity(nnz);
        let mut data: Vec<Complex64> = Vec::with_capacity(nnz);
        check(unsafe {
            qr_build_host(self.raw, row_lo, row_hi, indptr.as_mut_ptr(), indices.as_mut_ptr(),
                          data.as_mut_ptr() as *mut f64, QR_INDPTR_LOCAL)
        })?;
        unsafe { indptr.set_len(rows + 1); indices.set_len(nnz); data.set_len(nnz); }
        Ok(UnsafeVectors { data, indices, indptr })
    }

    /// The build followed by `util::csmatrix_eliminate_zeroes(m, tol)` (util.rs:154-171) as one device
    /// pass: counts + prefix scan give indptr and nnz, then only the kept entries are written.
    pub fn build_rows_compact(&self, row_lo: u64, row_hi: u64, tol: f64) -> Result<UnsafeVectors, QrustyErr> {
        let rows = (row_hi - row_lo) as usize;
        let null = std::ptr::null_mut::<c_void>();
        let (mut d_ptr, mut d_idx, mut d_dat) = (null, null, null);
        let mut nnz: u64 = 0;
        check(unsafe { qr_malloc_device(&mut d_ptr, (rows + 1) * 8) })?;
        let res = (|| {
            check(unsafe { qr_build_compact_count(self.raw, row_lo, row_hi, tol, d_ptr as *mut u64, &mut nnz, null) })?;
            let n = nnz as usize;
            check(unsafe { qr_malloc_device(&mut d_idx, n.max(2) * 8) })?;
            check(unsafe { qr_malloc_device(&mut d_dat, n.max(1) * 16) })?;
            check(unsafe { qr_build_compact_fill(self.raw, row_lo, row_hi, tol, d_ptr as *const u64,
                                                 d_idx as *mut u64, d_dat as *mut f64, null) })?;
            let mut indptr: Vec<u64> = Vec::with_capacity(rows + 1);
            let mut indices: Vec<u64> = Vec::with_capacity(n);
            let mut data: Vec<Complex64> = Vec::with_capacity(n);
            check(unsafe { qr_memcpy_d2h(indptr.as_mut_ptr() as *mut c_void, d_ptr, (rows + 1) * 8, null) })?;
            check(unsafe { qr_memcpy_d2h(indices.as_mut_ptr() as *mut c_void, d_idx, n * 8, null) })?;
            check(unsafe { qr_memcpy_d2h(data.as_mut_ptr() as *mut c_void, d_dat, n * 16, null) })?;
            unsafe { indptr.set_len(rows + 1); indices.set_len(n); data.set_len(n); }
            Ok(UnsafeVectors { data, indices, indptr })
        })();
        unsafe { qr_free_device(d_ptr); qr_free_device(d_idx); qr_free_device(d_dat); }
        res
    }

    /// `rawio::write` (rawio.rs:128-148) of rows [row_lo,row_hi), streamed from the GPU.
    pub fn write_rawio(&self, row_lo: u64, row_hi: u64, path: &std::path::Path) -> Result<(), QrustyErr> {
        let c = std::ffi::CString::new(path.to_string_lossy().as_bytes()).map_err(|_| QrustyErr::new("bad path"))?;
        check(unsafe { qr_write_rawio(self.raw, row_lo, row_hi, c.as_ptr()) })
    }

    /// Matrix-free y = H v (replaces build + `spmat_dot_densevec`, accel.rs:338-370).
    pub fn apply(&self, v: &[Complex64]) -> Result<Vec<Complex64>, QrustyErr> {
        assert_eq!(v.len() as u64, self.info.dim);
        let mut y: Vec<Complex64> = Vec::with_capacity(v.len());
        check(unsafe { qr_apply_host(self.raw, v.as_ptr() as *const f64, y.as_mut_ptr() as *mut f64) })?;
        unsafe { y.set_len(v.len()); }
        Ok(y)
    }
}

impl Drop for Plan {
    fn drop(&mut self) { unsafe { qr_plan_destroy(self.raw); } }
}

/// `precond2` (pyqrusty/src/lib.rs:457-468): dx / reg(diag - e, tol), host slices in and out.
pub fn precond2(diag: &[Complex64], dx: &[Complex64], e: Complex64, tol: f64) -> Result<Vec<Complex64>, QrustyErr> {
    assert_eq!(diag.len(), dx.len());
    let (n, bytes, null) = (dx.len(), dx.len() * 16, std::ptr::null_mut::<c_void>());
    let (mut d_diag, mut d_dx, mut d_out) = (null, null, null);
    let res = (|| {
        check(unsafe { qr_malloc_device(&mut d_diag, bytes.max(16)) })?;
        check(unsafe { qr_malloc_device(&mut d_dx, bytes.max(16)) })?;
        check(unsafe { qr_malloc_device(&mut d_out, bytes.max(16)) })?;
        check(unsafe { qr_memcpy_h2d(d_diag, diag.as_ptr() as *const c_void, bytes, null) })?;
        check(unsafe { qr_memcpy_h2d(d_dx, dx.as_ptr() as *const c_void, bytes, null) })?;
        let ev = [e.re, e.im];
        check(unsafe { qr_precond2_device(n as u64, d_diag as *const f64, d_dx as *const f64, ev.as_ptr(), tol,
                                          d_out as *mut f64, null) })?;
        let mut out: Vec<Complex64> = Vec::with_capacity(n);
        check(unsafe { qr_memcpy_d2h(out.as_mut_ptr() as *mut c_void, d_out, bytes, null) })?;
        unsafe { out.set_len(n); }
        Ok(out)
    })();
    unsafe { qr_free_device(d_diag); qr_free_device(d_dx); qr_free_device(d_out); }
    res
}

pub fn device_count() -> i32 {
    let mut n: c_int = 0;
    if unsafe { qr_device_count(&mut n) } == 0 { n } else { 0 }
}

/// Entry point used by `SparsePauliOp::to_matrix_cuda`.  `n_gpus` > 1
/// builds contiguous row blocks on devices 0..n_gpus and concatenates them on the host
/// (indptr rebased), which is what `AccelMode::Cuda(n)` means for a host-resident CsMatI.
pub fn make_unsafe_vectors_cuda(members: &[PauliSummand], n_gpus: usize) -> Result<UnsafeVectors, QrustyErr> {
    let n_gpus = n_gpus.max(1);
    if n_gpus == 1 {
        let plan = Plan::new(members, 0)?;
        return plan.build_rows(0, plan.info.dim);
    }
    let dim = 1u64 << members[0].0.num_qubits();
    let per = dim / n_gpus as u64;
    let parts: Vec<Result<UnsafeVectors, QrustyErr>> = std::thread::scope(|s| {
        let hs: Vec<_> = (0..n_gpus).map(|d| s.spawn(move || {
            let plan = Plan::new(members, d as i32)?;
            plan.build_rows(per * d as u64, per * (d as u64 + 1))
        })).collect();
        hs.into_iter().map(|h| h.join().unwrap()).collect()
    });
    let mut out = UnsafeVectors { data: Vec::new(), indices: Vec::new(), indptr: vec![0] };
    for part in parts {
        let part = part?;
        let base = *out.indptr.last().unwrap();
        out.indptr.extend(part.indptr[1..].iter().map(|v| v + base));
        out.indices.extend(part.indices);
        out.data.extend(part.data);
    }
    Ok(out)
}
