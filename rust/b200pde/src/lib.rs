//! Safe wrappers over `libb200pde.so`: the GPU side of rustpde's per-timestep spectral path behind the reference's
//! own trait surface.
//!
//! * [`GpuNavier2D`] implements `Integrate` (src/lib.rs:167-178): `update()` replaces `Navier2D::update()`
//!   (src/navier_stokes/navier.rs:438-466); state stays on the GPU between steps.
//! * [`GpuHholtzAdi`], [`GpuPoisson`], [`GpuHholtz`] implement `Solve<f64, Ix2>` (src/solver.rs:59-82): `input` is copied to
//!   the device, the solve runs there, `output` is copied back; `axis` is ignored exactly as in the reference
//!   (`#[allow(unused_variables)]`, hholtz_adi.rs:120).  Shape mismatches panic like `fdma_tensor.rs:256-263`.
//! * [`GpuField2`] mirrors `FieldBase::{forward, backward, to_ortho, from_ortho, gradient}` (src/field.rs:103-129).
//!
//! With the feature `rustpde-traits` the impls are for rustpde's own `Integrate` / `Solve` traits (build inside the
//! rustpde workspace); otherwise for the identical local copies below, so the crate also builds on its own.
//! NOTE: this image has no Rust toolchain (probed: no cargo / rustc), so these sources are unbuilt here; the same ABI
//! is exercised through ctypes (rustpde_mpi_b200/api.py) and through examples/cpp_driver.
use b200pde_sys as sys;
use ndarray::{Array1, Array2, ArrayBase, Data, DataMut, Ix2};
use std::ffi::{CStr, CString};
use std::os::raw::c_void;
use std::ptr;

#[cfg(feature = "rustpde-traits")]
pub use rustpde::{solver::Solve, Integrate};

/// src/lib.rs:167-178
#[cfg(not(feature = "rustpde-traits"))]
pub trait Integrate {
    fn update(&mut self);
    fn get_time(&self) -> f64;
    fn get_dt(&self) -> f64;
    fn callback(&mut self);
    fn exit(&mut self) -> bool;
}

/// src/solver.rs:59-82
#[cfg(not(feature = "rustpde-traits"))]
pub trait Solve<A, D> {
    fn solve<S1, S2>(&self, input: &ArrayBase<S1, D>, output: &mut ArrayBase<S2, D>, axis: usize)
    where
        S1: Data<Elem = A>,
        S2: Data<Elem = A> + DataMut;
    fn solve_par<S1, S2>(&self, input: &ArrayBase<S1, D>, output: &mut ArrayBase<S2, D>, axis: usize)
    where
        S1: Data<Elem = A>,
        S2: Data<Elem = A> + DataMut;
}

/// The reference panics on errors (shape mismatches, allocation): keep that behaviour.
fn check(status: i32) {
    if status != sys::B2_OK {
        let msg = unsafe { CStr::from_ptr(sys::b2_last_error()) }.to_string_lossy().into_owned();
        panic!("b200pde error {}: {}", status, msg);
    }
}

/// `BaseKind` (src/field.rs:173-177) + size, as `cheb_dirichlet(n)` etc. produce them
#[derive(Clone, Copy, Debug, PartialEq)]
pub enum Base {
    Chebyshev(usize),
    ChebDirichlet(usize),
    ChebNeumann(usize),
    ChebDirichletNeumann(usize),
    FourierR2c(usize),
}
impl Base {
    fn kind_n(self) -> (i32, i32) {
        match self {
            Base::Chebyshev(n) => (sys::B2_CHEBYSHEV, n as i32),
            Base::ChebDirichlet(n) => (sys::B2_CHEB_DIRICHLET, n as i32),
            Base::ChebNeumann(n) => (sys::B2_CHEB_NEUMANN, n as i32),
            Base::ChebDirichletNeumann(n) => (sys::B2_CHEB_DIRICHLET_NEUMANN, n as i32),
            Base::FourierR2c(n) => (sys::B2_FOURIER_R2C, n as i32),
        }
    }
}

/// One per GPU / MPI rank (replaces funspace `initialize()` -> `Universe`, src/mpi/mod.rs)
pub struct Context {
    raw: *mut sys::b2_ctx,
}
impl Context {
    pub fn new(device: i32) -> Self {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::b2_ctx_create(device, 0, 1, 0, &mut raw) });
        Context { raw }
    }
    /// One rank per GPU.  `exchange` must all-gather 64 bytes per rank in rank order (e.g. `MPI_Allgather`).
    pub fn new_distributed(device: i32, rank: i32, nranks: i32, heap_bytes: usize, exchange: impl Fn(&[u8; 64]) -> Vec<u8>) -> Self {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::b2_ctx_create(device, rank, nranks, heap_bytes, &mut raw) });
        let mut handle = [0u8; 64];
        check(unsafe { sys::b2_ctx_heap_handle(raw, handle.as_mut_ptr() as *mut c_void) });
        let all = exchange(&handle);
        assert_eq!(all.len(), 64 * nranks as usize);
        check(unsafe { sys::b2_ctx_attach_peers(raw, all.as_ptr() as *const c_void) });
        Context { raw }
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::b2_ctx_destroy(self.raw) };
    }
}

/// `Space2::new(&base0, &base1)` + `Field2::new(&space)` (src/field.rs:81-90): `v`, `vhat` live on the device
pub struct GpuField2 {
    space: *mut sys::b2_space,
    raw: *mut sys::b2_field,
    owned: bool,
}
impl GpuField2 {
    pub fn new(ctx: &Context, base0: Base, base1: Base) -> Self {
        let (k0, n0) = base0.kind_n();
        let (k1, n1) = base1.kind_n();
        let mut space = ptr::null_mut();
        check(unsafe { sys::b2_space2_create(ctx.raw, k0, n0, k1, n1, &mut space) });
        let mut raw = ptr::null_mut();
        check(unsafe { sys::b2_field_create(space, &mut raw) });
        GpuField2 { space, raw, owned: true }
    }
    fn shape(&self, kind: i32) -> (usize, usize, bool) {
        let (mut r, mut c, mut cx) = (0, 0, 0);
        check(unsafe { sys::b2_space_shape(self.space, kind, &mut r, &mut c, &mut cx) });
        (r as usize, c as usize, cx != 0)
    }
    /// `field.v.assign(&v)` (physical values, standard layout)
    pub fn set_v<S: Data<Elem = f64>>(&mut self, v: &ArrayBase<S, Ix2>) {
        let v = v.as_standard_layout();
        check(unsafe { sys::b2_field_set_v_host(self.raw, v.as_ptr() as *const c_void, v.len() * 8) });
    }
    pub fn v(&self) -> Array2<f64> {
        let (r, c, _) = self.shape(sys::B2_SHAPE_PHYSICAL);
        let mut out = Array2::<f64>::zeros((r, c));
        check(unsafe { sys::b2_field_get_v_host(self.raw, out.as_mut_ptr() as *mut c_void, out.len() * 8) });
        out
    }
    /// real spectral coefficients (Chebyshev x Chebyshev spaces)
    pub fn set_vhat<S: Data<Elem = f64>>(&mut self, vhat: &ArrayBase<S, Ix2>) {
        let a = vhat.as_standard_layout();
        check(unsafe { sys::b2_field_set_vhat_host(self.raw, a.as_ptr() as *const c_void, a.len() * 8) });
    }
    pub fn vhat(&self) -> Array2<f64> {
        let (r, c, cx) = self.shape(sys::B2_SHAPE_SPECTRAL);
        assert!(!cx, "complex spectral space: use vhat_complex");
        let mut out = Array2::<f64>::zeros((r, c));
        check(unsafe { sys::b2_field_get_vhat_host(self.raw, out.as_mut_ptr() as *mut c_void, out.len() * 8) });
        out
    }
    /// complex spectral coefficients (Fourier x Chebyshev): `Complex<f64>` is the interleaved (re, im) pair the ABI expects
    pub fn set_vhat_complex<S: Data<Elem = num_complex::Complex<f64>>>(&mut self, vhat: &ArrayBase<S, Ix2>) {
        let a = vhat.as_standard_layout();
        check(unsafe { sys::b2_field_set_vhat_host(self.raw, a.as_ptr() as *const c_void, a.len() * 16) });
    }
    pub fn vhat_complex(&self) -> Array2<num_complex::Complex<f64>> {
        let (r, c, cx) = self.shape(sys::B2_SHAPE_SPECTRAL);
        assert!(cx);
        let mut out = Array2::<num_complex::Complex<f64>>::zeros((r, c));
        check(unsafe { sys::b2_field_get_vhat_host(self.raw, out.as_mut_ptr() as *mut c_void, out.len() * 16) });
        out
    }
    /// src/field.rs:103-105
    pub fn forward(&mut self) {
        check(unsafe { sys::b2_forward(self.raw) });
    }
    /// src/field.rs:108-110
    pub fn backward(&mut self) {
        check(unsafe { sys::b2_backward(self.raw) });
    }
    /// src/field.rs:113-115 (real spaces)
    pub fn to_ortho(&self) -> Array2<f64> {
        let (r, c, cx) = self.shape(sys::B2_SHAPE_ORTHO);
        assert!(!cx);
        let mut arr = ptr::null_mut();
        check(unsafe { sys::b2_array_create(self.space, sys::B2_SHAPE_ORTHO, &mut arr) });
        check(unsafe { sys::b2_to_ortho(self.raw, arr) });
        let mut out = Array2::<f64>::zeros((r, c));
        check(unsafe { sys::b2_array_get_host(arr, out.as_mut_ptr() as *mut c_void, out.len() * 8) });
        unsafe { sys::b2_array_destroy(arr) };
        out
    }
    /// src/field.rs:127-129
    pub fn gradient(&self, deriv: [usize; 2], scale: Option<[f64; 2]>) -> Array2<f64> {
        let (r, c, cx) = self.shape(sys::B2_SHAPE_ORTHO);
        assert!(!cx);
        let mut arr = ptr::null_mut();
        check(unsafe { sys::b2_array_create(self.space, sys::B2_SHAPE_ORTHO, &mut arr) });
        let sc = scale.as_ref().map_or(ptr::null(), |s| s.as_ptr());
        check(unsafe { sys::b2_gradient(self.raw, deriv[0] as i32, deriv[1] as i32, sc, arr) });
        let mut out = Array2::<f64>::zeros((r, c));
        check(unsafe { sys::b2_array_get_host(arr, out.as_mut_ptr() as *mut c_void, out.len() * 8) });
        unsafe { sys::b2_array_destroy(arr) };
        out
    }
    /// `FieldBase::average_axis` (src/field/average.rs:26-35) on one rank: the dx-weighted mean of `v` along `axis`, reduced on
    /// the device (`w0`, `w1` = `dx / length` per axis, as the reference computes them from `self.dx`, `self.x`).  With several
    /// ranks pass the weights of this rank's rows as `w0` and combine the results like src/field_mpi/average.rs:15-61
    /// (axis 0: `all_gather_sum`; axis 1: concatenate the ranks' parts).
    pub fn average_axis(&self, axis: usize, w0_local: &[f64], w1: &[f64]) -> Array1<f64> {
        assert!(axis < 2);
        let mut arr = ptr::null_mut();
        check(unsafe { sys::b2_field_array(self.raw, 0, &mut arr) });
        let mut out = Array1::<f64>::zeros(if axis == 0 { w1.len() } else { w0_local.len() });
        let mode = if axis == 0 { 1 } else { 2 };
        check(unsafe { sys::b2_array_weighted_sum(arr, w0_local.as_ptr(), w1.as_ptr(), mode, out.as_mut_ptr()) });
        out
    }
    /// `FieldBase::average` (src/field/average.rs:53-59): this rank's part of the volume-weighted mean
    pub fn average(&self, w0_local: &[f64], w1: &[f64]) -> f64 {
        let mut arr = ptr::null_mut();
        check(unsafe { sys::b2_field_array(self.raw, 0, &mut arr) });
        let mut out = 0.0;
        check(unsafe { sys::b2_array_weighted_sum(arr, w0_local.as_ptr(), w1.as_ptr(), 0, &mut out) });
        out
    }
}
impl Drop for GpuField2 {
    fn drop(&mut self) {
        if self.owned {
            unsafe {
                sys::b2_field_destroy(self.raw);
                sys::b2_space_destroy(self.space);
            }
        }
    }
}

/// Shared body of the three field solvers: copy in, `b2_solve`, copy out.
struct GpuSolver {
    raw: *mut sys::b2_solver,
    space: *mut sys::b2_space,
}
impl GpuSolver {
    fn run<S1: Data<Elem = f64>, S2: Data<Elem = f64> + DataMut>(&self, input: &ArrayBase<S1, Ix2>, output: &mut ArrayBase<S2, Ix2>) {
        let (mut ain, mut aout) = (ptr::null_mut(), ptr::null_mut());
        check(unsafe { sys::b2_array_create(self.space, sys::B2_SHAPE_ORTHO, &mut ain) });
        check(unsafe { sys::b2_array_create(self.space, sys::B2_SHAPE_SPECTRAL, &mut aout) });
        let a = input.as_standard_layout();
        check(unsafe { sys::b2_array_set_host(ain, a.as_ptr() as *const c_void, a.len() * 8) });   // B2_ERR_SHAPE -> panic: "Dimension mismatch"
        check(unsafe { sys::b2_solve(self.raw, ain, aout) });
        let mut tmp = Array2::<f64>::zeros(output.raw_dim());
        check(unsafe { sys::b2_array_get_host(aout, tmp.as_mut_ptr() as *mut c_void, tmp.len() * 8) });
        output.assign(&tmp);
        unsafe {
            sys::b2_array_destroy(ain);
            sys::b2_array_destroy(aout);
        }
    }
}
impl Drop for GpuSolver {
    fn drop(&mut self) {
        unsafe { sys::b2_solver_destroy(self.raw) };
    }
}

macro_rules! impl_solve {
    ($t:ident) => {
        impl Solve<f64, Ix2> for $t {
            #[allow(unused_variables)]
            fn solve<S1, S2>(&self, input: &ArrayBase<S1, Ix2>, output: &mut ArrayBase<S2, Ix2>, axis: usize)
            where
                S1: Data<Elem = f64>,
                S2: Data<Elem = f64> + DataMut,
            {
                self.0.run(input, output);
            }
            #[allow(unused_variables)]
            fn solve_par<S1, S2>(&self, input: &ArrayBase<S1, Ix2>, output: &mut ArrayBase<S2, Ix2>, axis: usize)
            where
                S1: Data<Elem = f64>,
                S2: Data<Elem = f64> + DataMut,
            {
                self.0.run(input, output);
            }
        }
    };
}

/// `HholtzAdi::new(&field, c)` (src/solver/hholtz_adi.rs:48-76)
pub struct GpuHholtzAdi(GpuSolver);
impl GpuHholtzAdi {
    pub fn new(field: &GpuField2, c: [f64; 2]) -> Self {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::b2_hholtz_adi_create(field.raw, c[0], c[1], &mut raw) });
        GpuHholtzAdi(GpuSolver { raw, space: field.space })
    }
}
impl_solve!(GpuHholtzAdi);

/// `Poisson::new(&field, c)` (src/solver/poisson.rs:54-94).  `eig` = what `FdmaTensor::from_matrix` computed on the host
/// (`lam` already shifted by the singularity rule, `fwd`, `bwd`: fdma_tensor.rs:117-129); `None` for a Fourier axis 0.
pub struct GpuPoisson(GpuSolver);
impl GpuPoisson {
    pub fn new(field: &GpuField2, c: [f64; 2], eig: Option<(&[f64], &[f64], &[f64])>) -> Self {
        let mut raw = ptr::null_mut();
        let (l, f, b) = eig.map_or((ptr::null(), ptr::null(), ptr::null()), |(l, f, b)| (l.as_ptr(), f.as_ptr(), b.as_ptr()));
        check(unsafe { sys::b2_poisson_create(field.raw, c[0], c[1], l, f, b, &mut raw) });
        GpuPoisson(GpuSolver { raw, space: field.space })
    }
}
impl_solve!(GpuPoisson);

/// `Hholtz::new(&field, c)` (src/solver/hholtz.rs:66-101)
pub struct GpuHholtz(GpuSolver);
impl GpuHholtz {
    pub fn new(field: &GpuField2, c: [f64; 2], eig: Option<(&[f64], &[f64], &[f64])>) -> Self {
        let mut raw = ptr::null_mut();
        let (l, f, b) = eig.map_or((ptr::null(), ptr::null(), ptr::null()), |(l, f, b)| (l.as_ptr(), f.as_ptr(), b.as_ptr()));
        check(unsafe { sys::b2_hholtz_create(field.raw, c[0], c[1], l, f, b, &mut raw) });
        GpuHholtz(GpuSolver { raw, space: field.space })
    }
}
impl_solve!(GpuHholtz);

/// `Navier2D::new_confined / new_periodic` (src/navier_stokes/navier.rs:215-308, 336-428) with the state on the GPU
pub struct GpuNavier2D {
    raw: *mut sys::b2_navier,
    dt: f64,
    /// called by `callback()` with the downloaded state (IO / statistics stay in the reference's host code)
    pub on_callback: Option<Box<dyn FnMut(&GpuNavier2D)>>,
}
impl GpuNavier2D {
    #[allow(clippy::too_many_arguments)]
    pub fn new(ctx: &Context, nx: usize, ny: usize, ra: f64, pr: f64, dt: f64, aspect: f64, bc: &str, periodic: bool,
               pois_eig: Option<(&[f64], &[f64], &[f64])>) -> Self {
        let bc = CString::new(bc).unwrap();
        let (l, f, b) = pois_eig.map_or((ptr::null(), ptr::null(), ptr::null()), |(l, f, b)| (l.as_ptr(), f.as_ptr(), b.as_ptr()));
        let mut raw = ptr::null_mut();
        check(unsafe { sys::b2_navier2d_create(ctx.raw, nx as i32, ny as i32, ra, pr, dt, aspect, bc.as_ptr(), periodic as i32, l, f, b, &mut raw) });
        GpuNavier2D { raw, dt, on_callback: None }
    }
    /// 0 temp, 1 velx, 2 vely, 3 pres, 4 pseu, 5 tempbc: a borrowed view (upload / download with `set_vhat` / `vhat`)
    pub fn field(&self, which: i32) -> GpuField2 {
        let mut f = ptr::null_mut();
        check(unsafe { sys::b2_navier_field(self.raw, which, &mut f) });
        GpuField2 { space: ptr::null_mut(), raw: f, owned: false }
    }
    pub fn div_norm(&self) -> f64 {
        let mut d = 0.0;
        check(unsafe { sys::b2_navier_div_norm(self.raw, &mut d) });
        d
    }
    pub fn set_time(&mut self, t: f64) {
        check(unsafe { sys::b2_navier_set_time(self.raw, t) });
    }
    /// navier.rs:185-187
    pub fn reset_time(&mut self) {
        self.set_time(0.0);
    }
}
impl Integrate for GpuNavier2D {
    /// replaces navier.rs:438-466
    fn update(&mut self) {
        check(unsafe { sys::b2_navier_update(self.raw, 1) });
    }
    fn get_time(&self) -> f64 {
        let mut t = 0.0;
        check(unsafe { sys::b2_navier_get_time(self.raw, &mut t) });
        t
    }
    fn get_dt(&self) -> f64 {
        self.dt
    }
    fn callback(&mut self) {
        if let Some(mut cb) = self.on_callback.take() {
            cb(self);
            self.on_callback = Some(cb);
        }
    }
    /// navier.rs:482-489: stop when the divergence is NaN
    fn exit(&mut self) -> bool {
        self.div_norm().is_nan()
    }
}
impl Drop for GpuNavier2D {
    fn drop(&mut self) {
        unsafe { sys::b2_navier_destroy(self.raw) };
    }
}
