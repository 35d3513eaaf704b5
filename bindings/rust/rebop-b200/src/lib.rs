//! Ensembles of rebop `Gillespie` problems on a GPU, behind the shapes of rebop's own API
//! (`gillespie::Gillespie::{new_with_seed, add_reaction, advance_until, get_species}`, `Rate::lma`,
//! `define_system!`).  Every trajectory reproduces what `rebop` computes on the CPU for the same seed,
//! bit for bit.
//!
//! NOTE: this crate was written without a Rust toolchain at hand (the build image of the engine has no
//! cargo/rustc); it follows the C header one to one but has not been compiled.  The same calls are
//! exercised from Python (`rebop_b200/_ffi.py`) and C (`examples/c_abi_demo.c`).
use rebop_b200_sys as sys;
use std::ffi::{CStr, CString};
use std::ptr;

/// Error of the engine: status code of `include/rebop_b200.h` and the library's message.
#[derive(Debug, Clone)]
pub struct Error {
    pub status: i32,
    pub message: String,
}

impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "rebop_b200 error {}: {}", self.status, self.message)
    }
}
impl std::error::Error for Error {}

fn check(status: i32) -> Result<(), Error> {
    if status == sys::REBOP_OK {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(sys::rebop_b200_last_error()) }.to_string_lossy().into_owned();
    Err(Error { status, message })
}

/// rebop panics on misuse of `add_reaction` (`src/gillespie.rs:227-237`); so does this wrapper.
fn check_or_panic(status: i32) {
    if let Err(e) = check(status) {
        panic!("{}", e.message);
    }
}

/// A reaction rate, as in `rebop::gillespie::Rate` (law of mass action forms).
pub enum Rate {
    /// `Rate::lma(k, exponents)`: dense vector of reactant exponents, one per species.
    Lma(f64, Vec<u32>),
    /// `Rate::lma_sparse(k, [(species, exponent), ..])`.
    LmaSparse(f64, Vec<(u32, u32)>),
    /// `Rate::expr(e)` lowered to the post-order program of `Expr::eval` (`src/expr.rs:24-38`).
    Expr(Vec<sys::rebop_expr_op>),
}

/// N independent `Gillespie` problems resident on one GPU.
pub struct GillespieBatch {
    net: *mut sys::rebop_network,
    batch: *mut sys::rebop_batch,
    nb_species: usize,
    n: usize,
    x0: Vec<i64>,
    seeds: Vec<u64>,
    device: i32,
}

// One thread at a time per handle, like `&mut self`; the handle may move between threads.
unsafe impl Send for GillespieBatch {}

impl GillespieBatch {
    /// N x `Gillespie::new_with_seed(species, _, seeds[n])` (`src/gillespie.rs:179-187`).
    pub fn new_with_seeds<V: AsRef<[isize]>>(species: V, seeds: &[u64], device: i32) -> Self {
        let x0: Vec<i64> = species.as_ref().iter().map(|&v| v as i64).collect();
        let mut net = ptr::null_mut();
        check_or_panic(unsafe { sys::rebop_network_create(x0.len() as u32, sys::REBOP_ARITH_API, &mut net) });
        GillespieBatch { net, batch: ptr::null_mut(), nb_species: x0.len(), n: seeds.len(), x0, seeds: seeds.to_vec(), device }
    }

    /// `nb_species()` / `nb_reactions()` (`src/gillespie.rs:200,210`).
    pub fn nb_species(&self) -> usize {
        self.nb_species
    }
    pub fn nb_reactions(&self) -> usize {
        let mut n = 0u32;
        check_or_panic(unsafe { sys::rebop_network_nb_reactions(self.net, &mut n) });
        n as usize
    }

    /// `add_reaction(rate, differences)` (`src/gillespie.rs:225-244`); must precede the first advance.
    pub fn add_reaction<V: AsRef<[isize]>>(&mut self, rate: Rate, differences: V) {
        assert!(self.batch.is_null(), "add_reaction after the ensemble was started");
        let d: Vec<i64> = differences.as_ref().iter().map(|&v| v as i64).collect();
        assert_eq!(d.len(), self.nb_species);
        let status = match rate {
            Rate::Lma(k, exponents) => {
                assert_eq!(exponents.len(), self.nb_species);
                unsafe { sys::rebop_network_add_reaction_lma(self.net, k, exponents.as_ptr(), d.as_ptr()) }
            }
            Rate::LmaSparse(k, terms) => {
                let (idx, ex): (Vec<u32>, Vec<u32>) = terms.into_iter().unzip();
                unsafe { sys::rebop_network_add_reaction_lma_sparse(self.net, k, idx.as_ptr(), ex.as_ptr(), idx.len(), d.as_ptr()) }
            }
            Rate::Expr(program) => unsafe {
                sys::rebop_network_add_reaction_expr(self.net, program.as_ptr(), program.len(), d.as_ptr())
            },
        };
        check_or_panic(status);
    }

    fn start(&mut self) -> Result<(), Error> {
        if self.batch.is_null() {
            check(unsafe {
                sys::rebop_batch_create(self.net, self.device, self.n, self.x0.as_ptr(), 0, self.seeds.as_ptr(), 0, &mut self.batch)
            })?;
        }
        Ok(())
    }

    /// `advance_until(tmax)` on every trajectory (`src/gillespie.rs:315-344`).
    pub fn advance_until(&mut self, tmax: f64) -> Result<(), Error> {
        self.start()?;
        check(unsafe { sys::rebop_batch_advance_until(self.batch, tmax) })
    }

    /// The binding's loop `for i in 0..=nb_steps { advance_until(tmax*i/nb_steps); record }`
    /// (`src/pyo3_gillespie.rs:197-208`) in one launch; samples as `[step][species][trajectory]`.
    pub fn run_grid(&mut self, tmax: f64, nb_steps: u32) -> Result<Vec<i64>, Error> {
        self.start()?;
        check(unsafe { sys::rebop_batch_run_grid(self.batch, tmax, nb_steps, ptr::null(), 0, ptr::null_mut()) })?;
        let mut out = vec![0i64; (nb_steps as usize + 1) * self.nb_species * self.n];
        check(unsafe { sys::rebop_batch_samples_host_i64(self.batch, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// The `nb_steps = 0` branch (`src/pyo3_gillespie.rs:209-223`): every reaction of every trajectory.
    /// Returns (offsets `[n+1]`, times `[rows]`, samples `[species][rows]`).
    pub fn run_events(&mut self, tmax: f64) -> Result<(Vec<u64>, Vec<f64>, Vec<i32>), Error> {
        self.start()?;
        check(unsafe { sys::rebop_batch_run_events(self.batch, tmax, ptr::null(), 0) })?;
        let (mut rows, mut n_save) = (0u64, 0u32);
        check(unsafe { sys::rebop_batch_events_log_size(self.batch, &mut rows, &mut n_save) })?;
        let mut offsets = vec![0u64; self.n + 1];
        let mut times = vec![0f64; rows as usize];
        let mut samples = vec![0i32; rows as usize * n_save as usize];
        check(unsafe { sys::rebop_batch_events_log_host(self.batch, offsets.as_mut_ptr(), times.as_mut_ptr(), samples.as_mut_ptr()) })?;
        Ok((offsets, times, samples))
    }

    /// `get_species` for every trajectory, `[trajectory][species]` (`src/gillespie.rs:255-259`).
    pub fn get_species(&mut self) -> Result<Vec<i64>, Error> {
        self.start()?;
        let mut out = vec![0i64; self.n * self.nb_species];
        check(unsafe { sys::rebop_batch_get_species(self.batch, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// `get_time` for every trajectory (`src/gillespie.rs:246-248`).
    pub fn get_time(&mut self) -> Result<Vec<f64>, Error> {
        self.start()?;
        let mut out = vec![0f64; self.n];
        check(unsafe { sys::rebop_batch_get_time(self.batch, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// `seed(u64)` for every trajectory (`src/gillespie.rs:189-191`).
    pub fn seed(&mut self, seeds: &[u64]) -> Result<(), Error> {
        assert_eq!(seeds.len(), self.n);
        self.seeds = seeds.to_vec();
        if self.batch.is_null() {
            return Ok(());
        }
        check(unsafe { sys::rebop_batch_seed(self.batch, seeds.as_ptr(), 0) })
    }

    /// Exact ensemble sums of the last `run_grid`: (sum x, sum x^2) per (step, species).
    pub fn sample_sums(&mut self, rows: usize) -> Result<(Vec<i64>, Vec<u64>), Error> {
        let (mut s1, mut s2) = (vec![0i64; rows], vec![0u64; rows]);
        check(unsafe { sys::rebop_batch_sample_sums(self.batch, s1.as_mut_ptr(), s2.as_mut_ptr()) })?;
        Ok((s1, s2))
    }

    /// Reactions applied since creation.
    pub fn events(&mut self) -> Result<u64, Error> {
        let mut total = 0u64;
        check(unsafe { sys::rebop_batch_events(self.batch, &mut total, ptr::null_mut()) })?;
        Ok(total)
    }
}

impl Drop for GillespieBatch {
    fn drop(&mut self) {
        unsafe {
            sys::rebop_batch_destroy(self.batch);
            sys::rebop_network_destroy(self.net);
        }
    }
}

/// An ensemble of the struct `define_system!` generates (`src/gillespie_macro.rs:62-126`), built from the
/// macro's own text (see INTEGRATION.md section 3 for the `stringify!` arm that produces it).
pub struct SystemBatch {
    sys_handle: *mut sys::rebop_system,
    net: *mut sys::rebop_network,
    batch: *mut sys::rebop_batch,
    n_species: usize,
    n_reactions: usize,
    n: usize,
}

unsafe impl Send for SystemBatch {}

impl SystemBatch {
    /// `Name::with_parameters(params..)`, species set to `x0`, trajectory i seeded with `seed + i`.
    pub fn new(dsl: &str, params: &[f64], x0: &[i64], n: usize, seed: u64, device: i32) -> Result<Self, Error> {
        let text = CString::new(dsl).expect("define_system text contains a NUL byte");
        let mut sys_handle = ptr::null_mut();
        check(unsafe { sys::rebop_system_parse(text.as_ptr(), &mut sys_handle) })?;
        let (mut np, mut ns, mut nr) = (0u32, 0u32, 0u32);
        check(unsafe { sys::rebop_system_counts(sys_handle, &mut np, &mut ns, &mut nr) })?;
        assert_eq!(x0.len(), ns as usize);
        let mut net = ptr::null_mut();
        let mut this = SystemBatch { sys_handle, net, batch: ptr::null_mut(), n_species: ns as usize, n_reactions: nr as usize, n };
        check(unsafe { sys::rebop_system_network(sys_handle, params.as_ptr(), params.len(), &mut net) })?;
        this.net = net;
        check(unsafe { sys::rebop_batch_create(net, device, n, x0.as_ptr(), 0, ptr::null(), seed, &mut this.batch) })?;
        Ok(this)
    }

    /// `advance_until(tmax)` (`src/gillespie_macro.rs:98-126`).
    pub fn advance_until(&mut self, tmax: f64) -> Result<(), Error> {
        check(unsafe { sys::rebop_batch_advance_until(self.batch, tmax) })
    }

    /// The species fields of every trajectory, `[trajectory][species]` in the order of the macro's species list.
    pub fn species(&mut self) -> Result<Vec<i64>, Error> {
        let mut out = vec![0i64; self.n * self.n_species];
        check(unsafe { sys::rebop_batch_get_species(self.batch, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// Writing the parameter fields: the new values take effect at the next `advance_until`.
    pub fn set_parameters(&mut self, params: &[f64]) -> Result<(), Error> {
        let mut rates = vec![0f64; self.n_reactions];
        check(unsafe { sys::rebop_system_rates(self.sys_handle, params.as_ptr(), params.len(), rates.as_mut_ptr()) })?;
        check(unsafe { sys::rebop_batch_set_rates(self.batch, rates.as_ptr(), rates.len()) })
    }

    /// `seed(u64)`: trajectory i restarts its stream from `seed + i`.
    pub fn seed(&mut self, seed: u64) -> Result<(), Error> {
        check(unsafe { sys::rebop_batch_seed(self.batch, ptr::null(), seed) })
    }
}

impl Drop for SystemBatch {
    fn drop(&mut self) {
        unsafe {
            sys::rebop_batch_destroy(self.batch);
            sys::rebop_network_destroy(self.net);
            sys::rebop_system_destroy(self.sys_handle);
        }
    }
}

/// `define_system!` for GPU ensembles, with the grammar of rebop's macro (`src/gillespie_macro.rs:49-61`):
///
/// ```ignore
/// rebop_b200::define_system! {
///     r1 r2;
///     SIR { S, I, R }
///     r_infection: S + I  => I + I    @ r1
///     r_remission: I      => R        @ r2
/// }
/// let mut sir = SIR::with_parameters(1e-4, 0.01);
/// sir.S = 999;
/// sir.I = 1;
/// let mut ensemble = sir.ensemble(1_000_000, 0, 0)?;   // trajectory i is rebop's SIR seeded with i
/// ensemble.advance_until(250.0)?;
/// ```
///
/// rebop expands the system into straight-line Rust; here the text of the invocation is the system (the engine parses
/// the same grammar, `rebop_system_parse`), and the network-specialised CUDA kernel for it is compiled at build time
/// when the same text sits in `rebop-b200-sys/systems/<name>.rsys` (see that crate's `build.rs`: rebop_sysgen + nvcc),
/// or by NVRTC on first use otherwise.  The generated struct has the fields of rebop's (`src/gillespie_macro.rs:62-71`):
/// one `isize` per species -- the initial count of every trajectory -- and one `f64` per parameter.
#[macro_export]
macro_rules! define_system {
    ($($tt:tt)*) => {
        $crate::__define_system_impl! { text = { stringify!($($tt)*) }; $($tt)* }
    };
}

#[doc(hidden)]
#[macro_export]
macro_rules! __define_system_impl {
    (
      text = { $text:expr };
      $($param:ident)*;
      $name:ident { $($species:ident),* }
      $($rname:ident:
          $($($nr:literal)? $r:ident)? $(+ $($tnr:literal)? $tr:ident)* =>
          $($($np:literal)? $p:ident)? $(+ $($tnp:literal)? $tp:ident)*
          @ $rate:expr)*
    ) => {
        #[allow(non_snake_case)]
        #[derive(Clone, Debug)]
        struct $name {
            $(pub $species: isize,)*
            $(pub $param: f64,)*
        }
        #[allow(non_snake_case, dead_code)]
        impl $name {
            /// The text of the invocation: what the engine parses and what the build-time kernel is keyed by.
            pub const TEXT: &'static str = $text;
            /// `Name::new()` (`src/gillespie_macro.rs:75-82`): species 0, parameters NaN.
            fn new() -> Self {
                $name { $($species: 0,)* $($param: f64::NAN,)* }
            }
            /// `Name::with_parameters(..)` (`src/gillespie_macro.rs:91-98`).
            fn with_parameters($($param: f64),*) -> Self {
                $name { $($species: 0,)* $($param,)* }
            }
            /// `n` copies of this problem resident on GPU `device`; trajectory i is seeded with `seed + i`
            /// (`seed(u64)`, `src/gillespie_macro.rs:84-88`).
            fn ensemble(&self, n: usize, seed: u64, device: i32) -> Result<$crate::SystemBatch, $crate::Error> {
                $crate::SystemBatch::new(Self::TEXT, &[$(self.$param),*], &[$(self.$species as i64),*], n, seed, device)
            }
        }
    };
}
