//! UNBUILT IN THIS ENVIRONMENT (no Rust toolchain).  Safe façade over `incerto-cuda-sys` that keeps the method names
//! of the reference: `SimulationBuilder::{new, add_entity_spawner, add_systems, add_spatial_grid,
//! record_aggregate_time_series, record_time_series, build}` (src/simulation_builder.rs:39-306) and
//! `Simulation::{run, reset, count, sample, sample_single, sample_aggregate, get_aggregate_time_series,
//! get_time_series}` (src/simulation.rs:22-259).  The C++ façade (include/incerto_b200.hpp) and the Python mirror
//! (incerto_b200/simulation.py) implement the same mapping and are what the test-suite exercises.
//!
//! Differences from the reference, all forced by "systems are CUDA kernels":
//! * a `Component` struct of the reference becomes one engine component per numeric field, registered with
//!   [`SimulationBuilder::component`]; field-less structs become markers ([`SimulationBuilder::marker`]);
//! * `add_systems` takes values of [`System`] (the example systems) instead of arbitrary Rust functions;
//! * `With<T>` / `Without<T>` tuples become a [`Filter`].
use incerto_cuda_sys as sys;
use std::ffi::{c_void, CStr, CString};
use std::marker::PhantomData;

// ------------------------------------------------------------------------------------------------ errors
/// src/error.rs:12-43, rebuilt from the status code (same variants, same order).
pub use sys::SamplingError;
/// src/error.rs:47-52.
pub use sys::BuilderError;

/// Everything a call can report.  `Engine` carries the codes with no reference counterpart (the reference panics
/// there: sample_interval == 0, entity outside the grid bounds, NaN in an ordered aggregate, ...).
#[derive(Debug, Clone, PartialEq)]
pub enum Error {
    Sampling(SamplingError),
    Builder(BuilderError),
    Engine { code: i32, message: String },
}

fn check(code: i32) -> Result<(), Error> {
    use SamplingError::*;
    Err(match code {
        0 => return Ok(()),
        1 => Error::Sampling(ComponentDoesNotExist),
        2 => Error::Sampling(SingleNoEntities),
        3 => Error::Sampling(SingleMultipleEntities),
        4 => Error::Sampling(AggregateNoEntities),
        5 => Error::Sampling(TimeSeriesNotRecorded),
        6 => Error::Sampling(EntityIdentifierNotFound),
        7 => Error::Sampling(EntityIdentifierNotUnique),
        16 => Error::Builder(BuilderError::TimeSeriesRecordingConflict),
        code => Error::Engine {
            code,
            // SAFETY: the engine returns a NUL-terminated, thread-local string that lives until the next call
            message: unsafe { CStr::from_ptr(sys::incerto_last_error_message()) }.to_string_lossy().into_owned(),
        },
    })
}

// -------------------------------------------------------------------------------------------- components
/// Storage types of the device columns (`incerto_dtype`, include/incerto_b200.h).
pub trait Scalar: Copy + 'static {
    const DTYPE: i32;
}
macro_rules! scalar { ($($t:ty => $d:expr),*) => { $(impl Scalar for $t { const DTYPE: i32 = $d; })* } }
scalar!(u8 => 1, u16 => 2, u32 => 3, u64 => 4, i8 => 5, i16 => 6, i32 => 7, i64 => 8, f32 => 9, f64 => 10);

/// Handle of a registered component column: `LANES` values of `T` per entity (2 / 3 for `GridPosition<IVec2/IVec3>`).
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct Component<T, const LANES: usize = 1> {
    handle: sys::incerto_component,
    _t: PhantomData<T>,
}
/// Handle of a field-less component (`Person`, `Infected`, `GroupA` ...): one bit per entity.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct Marker {
    handle: sys::incerto_component,
}

/// `(With<A>, With<B>, Without<C>)` of the reference's query filters.
#[derive(Debug, Clone, Default)]
pub struct Filter {
    with: Vec<sys::incerto_component>,
    without: Vec<sys::incerto_component>,
}
impl Filter {
    pub fn all() -> Self { Self::default() }
    pub fn with(mut self, m: Marker) -> Self { self.with.push(m.handle); self }
    pub fn without(mut self, m: Marker) -> Self { self.without.push(m.handle); self }
    fn raw(&self) -> sys::incerto_filter {
        sys::incerto_filter { with: self.with.as_ptr(), n_with: self.with.len() as u32,
                              without: self.without.as_ptr(), n_without: self.without.len() as u32 }
    }
}

// --------------------------------------------------------------------------------------------- spawning
/// What a spawner closure receives (src/spawner.rs:3-12).  Entities are staged per archetype in SoA form and shipped
/// in bulk; spawn order = uid order is kept.
#[derive(Default)]
pub struct Spawner {
    batches: Vec<Batch>,
}
struct Batch {
    n: u64,
    columns: Vec<(sys::incerto_component, Vec<u8>)>, // raw little-endian column bytes; empty = marker on every row
}
impl Spawner {
    /// `n` entities of one archetype; `columns` built with [`column`] / [`marker_column`].
    pub fn spawn_batch(&mut self, n: u64, columns: Vec<(sys::incerto_component, Vec<u8>)>) {
        self.batches.push(Batch { n, columns });
    }
}
/// One data column of a batch (`values.len()` = n * LANES).
pub fn column<T: Scalar, const L: usize>(c: Component<T, L>, values: &[T]) -> (sys::incerto_component, Vec<u8>) {
    // SAFETY: T is a plain scalar; the slice is reinterpreted as its bytes
    let bytes = unsafe { std::slice::from_raw_parts(values.as_ptr() as *const u8, std::mem::size_of_val(values)) };
    (c.handle, bytes.to_vec())
}
/// A marker carried by every entity of the batch.
pub fn marker_column(m: Marker) -> (sys::incerto_component, Vec<u8>) { (m.handle, Vec::new()) }

// ---------------------------------------------------------------------------------------------- systems
/// The systems the engine has kernels for (`incerto_system_kind`); parameters as in the example programs.
#[derive(Debug, Clone)]
pub enum System {
    /// examples/coin_toss.rs:67-81
    CoinToss { num_heads: Component<u32>, num_tosses: Component<u32>, odds_heads: f64 },
    /// examples/forest_fire.rs:282-329
    FireSpread { position: Component<i16, 2>, cell: Component<u8>, probability: f64, burn_duration: u32 },
    /// examples/forest_fire.rs:332-351
    BurnProgression { position: Component<i16, 2>, cell: Component<u8> },
    /// examples/forest_fire.rs:354-365
    Regrowth { position: Component<i16, 2>, cell: Component<u8>, probability: f64 },
    /// Any other kind with its `repr(C)` parameter block from include/incerto_b200.h (pandemic, pandemic_spatial,
    /// traders, air_traffic_3d: the blocks are declared in incerto-cuda-sys next to the functions).
    Raw { kind: u32, params: Vec<u8> },
}

#[repr(C)] struct CoinTossParams { num_heads: i32, num_tosses: i32, odds_heads: f64 }
#[repr(C)] struct ForestParams { position: i32, cell: i32, probability: f64, burn_duration: u32 }

fn bytes_of<P>(p: &P) -> Vec<u8> {
    // SAFETY: P is repr(C) plain data
    unsafe { std::slice::from_raw_parts(p as *const P as *const u8, std::mem::size_of::<P>()) }.to_vec()
}
impl System {
    fn lower(&self) -> (u32, Vec<u8>) {
        match self {
            System::CoinToss { num_heads, num_tosses, odds_heads } =>
                (10, bytes_of(&CoinTossParams { num_heads: num_heads.handle, num_tosses: num_tosses.handle, odds_heads: *odds_heads })),
            System::FireSpread { position, cell, probability, burn_duration } =>
                (20, bytes_of(&ForestParams { position: position.handle, cell: cell.handle, probability: *probability, burn_duration: *burn_duration })),
            System::BurnProgression { position, cell } =>
                (21, bytes_of(&ForestParams { position: position.handle, cell: cell.handle, probability: 0.0, burn_duration: 0 })),
            System::Regrowth { position, cell, probability } =>
                (22, bytes_of(&ForestParams { position: position.handle, cell: cell.handle, probability: *probability, burn_duration: 0 })),
            System::Raw { kind, params } => (*kind, params.clone()),
        }
    }
}

// ------------------------------------------------------------------------------------------- aggregates
/// `SampleAggregate` implementations of the reference (src/util/aggregate.rs:73-192 and the examples').
#[derive(Debug, Clone)]
pub struct Aggregate {
    kind: u32,
    component: sys::incerto_component,
    component2: sys::incerto_component,
    percentile: u32,
    filter: Filter,
}
impl Aggregate {
    fn of(kind: u32, component: i32) -> Self { Self { kind, component, component2: -1, percentile: 0, filter: Filter::all() } }
    pub fn minimum<T: Scalar>(c: Component<T>) -> Self { Self::of(3, c.handle) }
    pub fn maximum<T: Scalar>(c: Component<T>) -> Self { Self::of(4, c.handle) }
    pub fn mean<T: Scalar>(c: Component<T>) -> Self { Self::of(5, c.handle) }
    pub fn median<T: Scalar>(c: Component<T>) -> Self { Self::of(6, c.handle) }
    pub fn percentile<T: Scalar>(c: Component<T>, p: u32) -> Self { Self { percentile: p, ..Self::of(7, c.handle) } }
    /// examples/coin_toss.rs:33-49 (f64)
    pub fn coin_odds(num_heads: Component<u32>, num_tosses: Component<u32>) -> Self { Self { component2: num_tosses.handle, ..Self::of(16, num_heads.handle) } }
    /// examples/forest_fire.rs:104-134 ([`FireStats`])
    pub fn fire_stats(cell: Component<u8>) -> Self { Self::of(17, cell.handle) }
    pub fn filtered(mut self, f: Filter) -> Self { self.filter = f; self }
    fn raw(&self) -> sys::incerto_aggregate_desc {
        sys::incerto_aggregate_desc { component: self.component, component2: self.component2, kind: self.kind,
                                      percentile: self.percentile, param: 0, filter: self.filter.raw() }
    }
}
/// examples/forest_fire.rs:94-101
#[repr(C)]
#[derive(Debug, Clone, Copy, Default, PartialEq, Eq)]
pub struct FireStats { pub healthy_count: u64, pub burning_count: u64, pub burned_count: u64, pub empty_count: u64, pub total_cells: u64 }

/// src/types/times_series.rs:1-6: borrowed from the simulation, valid until its next `run` / `reset`.
pub struct TimeSeries<'a, T> {
    pub time: &'a [u64],
    pub values: &'a [T],
    pub sample_interval: u64,
}

// ---------------------------------------------------------------------------------------------- builder
/// src/simulation_builder.rs:25-306.  Methods take `self` by value like the reference's.
pub struct SimulationBuilder {
    raw: *mut sys::incerto_builder,
    spawners: Vec<Box<dyn FnOnce(&mut Spawner)>>,
    first_error: Option<Error>,
}
impl SimulationBuilder {
    /// `SimulationBuilder::new` (:42-56)
    pub fn new() -> Self {
        let mut raw = std::ptr::null_mut();
        let first_error = check(unsafe { sys::incerto_builder_new(&mut raw) }).err();
        Self { raw, spawners: Vec::new(), first_error }
    }
    fn note(mut self, r: Result<(), Error>) -> Self {
        if self.first_error.is_none() { self.first_error = r.err(); }
        self
    }
    /// Engine knob (the reference seeds from the OS): the Philox key of every draw, DESIGN.md §2.
    pub fn seed(self, seed: u64, replica: u32) -> Self {
        let r = check(unsafe { sys::incerto_builder_set_seed(self.raw, seed, replica) });
        self.note(r)
    }
    /// One numeric field of a reference `Component` struct.
    pub fn component<T: Scalar, const L: usize>(&mut self, name: &str) -> Component<T, L> {
        let mut handle = -1;
        let name = CString::new(name).expect("component name");
        let r = check(unsafe { sys::incerto_builder_register_component(self.raw, name.as_ptr(), T::DTYPE, L as u32, &mut handle) });
        if self.first_error.is_none() { self.first_error = r.err(); }
        Component { handle, _t: PhantomData }
    }
    /// A field-less component.
    pub fn marker(&mut self, name: &str) -> Marker {
        let mut handle = -1;
        let name = CString::new(name).expect("component name");
        let r = check(unsafe { sys::incerto_builder_register_component(self.raw, name.as_ptr(), 0, 1, &mut handle) });
        if self.first_error.is_none() { self.first_error = r.err(); }
        Marker { handle }
    }
    /// `add_entity_spawner` (:154-158): the closures run once, in insertion order, inside `build` (:295-305).
    pub fn add_entity_spawner(mut self, f: impl FnOnce(&mut Spawner) + 'static) -> Self {
        self.spawners.push(Box::new(f));
        self
    }
    /// `add_systems` (:66-77)
    pub fn add_systems(self, systems: &[System]) -> Self {
        let lowered: Vec<(u32, Vec<u8>)> = systems.iter().map(System::lower).collect();
        let descs: Vec<sys::incerto_system_desc> = lowered.iter()
            .map(|(kind, p)| sys::incerto_system_desc { kind: *kind, params: p.as_ptr() as *const c_void, params_size: p.len() as u32 })
            .collect();
        let r = check(unsafe { sys::incerto_builder_add_systems(self.raw, descs.as_ptr(), descs.len() as u32) });
        self.note(r)
    }
    /// `add_spatial_grid::<IVec2, C>` (:114-139); `bounds` = (min, max) inclusive or `None` for an unbounded grid.
    pub fn add_spatial_grid_2d<T: Scalar>(self, position: Component<i16, 2>, component: Component<T>, bounds: Option<([i32; 2], [i32; 2])>) -> Self {
        let flat = bounds.map(|(lo, hi)| [lo[0], lo[1], hi[0], hi[1]]);
        let mut grid = -1;
        let r = check(unsafe {
            sys::incerto_builder_add_spatial_grid(self.raw, 2, flat.as_ref().map_or(std::ptr::null(), |b| b.as_ptr()),
                                                  position.handle, component.handle, &mut grid)
        });
        self.note(r)
    }
    /// `record_aggregate_time_series` (:262-287); `sample_interval == 0` is the reference's `assert!` (:271).
    pub fn record_aggregate_time_series(self, aggregate: &Aggregate, sample_interval: u64) -> Self {
        let r = check(unsafe { sys::incerto_builder_record_aggregate_time_series(self.raw, &aggregate.raw(), sample_interval) });
        self.note(r)
    }
    /// `record_time_series::<C, I>` (:185-206)
    pub fn record_time_series<T: Scalar, I: Scalar>(self, component: Component<T>, identifier: Component<I>, sample_interval: u64) -> Self {
        let r = check(unsafe { sys::incerto_builder_record_time_series(self.raw, component.handle, identifier.handle, sample_interval) });
        self.note(r)
    }
    /// `build` (:295-305): runs the spawners and hands the entities to the device.
    pub fn build(mut self) -> Result<Simulation, Error> {
        if let Some(e) = self.first_error.take() { return Err(e); }
        // the staging buffers outlive incerto_builder_build, so they can be passed without a copy
        let mut staged = Spawner::default();
        for f in std::mem::take(&mut self.spawners) { f(&mut staged); }
        for batch in &staged.batches {
            let cols: Vec<sys::incerto_column_data> = batch.columns.iter()
                .map(|(c, bytes)| sys::incerto_column_data { component: *c, data: if bytes.is_empty() { std::ptr::null() } else { bytes.as_ptr() as *const c_void } })
                .collect();
            check(unsafe { sys::incerto_builder_add_entity_spawner_borrowed(self.raw, batch.n, cols.as_ptr(), cols.len() as u32) })?;
        }
        let mut sim = std::ptr::null_mut();
        let r = check(unsafe { sys::incerto_builder_build(self.raw, &mut sim) });
        self.raw = std::ptr::null_mut(); // consumed by build, successful or not
        r.map(|()| Simulation { raw: sim })
    }
}
impl Drop for SimulationBuilder {
    fn drop(&mut self) {
        if !self.raw.is_null() { unsafe { sys::incerto_builder_free(self.raw) } }
    }
}

// ------------------------------------------------------------------------------------------- simulation
/// src/simulation.rs:17-259.  One handle = one host thread = one device.
pub struct Simulation {
    raw: *mut sys::incerto_sim,
}
impl Simulation {
    /// `run(num_steps)` (:25-31): returns after the device finished.
    pub fn run(&mut self, num_steps: usize) -> Result<(), Error> { check(unsafe { sys::incerto_sim_run(self.raw, num_steps as u64) }) }
    /// Restart from the spawners under another replica key, reusing all device memory (:22-24 of the builder docs).
    /// Not available after `build` passed borrowed staging buffers (this façade does): build a new simulation instead.
    pub fn reset(&mut self, replica: u32) -> Result<(), Error> { check(unsafe { sys::incerto_sim_reset(self.raw, replica) }) }
    /// `count::<F>()` (:163-173)
    pub fn count(&self, filter: &Filter) -> Result<usize, Error> {
        let mut n = 0u64;
        check(unsafe { sys::incerto_sim_count(self.raw, &filter.raw(), &mut n) })?;
        Ok(n as usize)
    }
    /// `sample_aggregate::<C, Out>()` (:107-150); `Out` must be the aggregate's output type (e.g. `f64` for
    /// `coin_odds`, [`FireStats`] for `fire_stats`, `T` for `minimum::<T>`).
    pub fn sample_aggregate<Out: Copy + Default>(&self, aggregate: &Aggregate) -> Result<Out, Error> {
        let mut out = Out::default();
        check(unsafe { sys::incerto_sim_sample_aggregate(self.raw, &aggregate.raw(), &mut out as *mut Out as *mut c_void, std::mem::size_of::<Out>()) })?;
        Ok(out)
    }
    /// `sample::<C, I>(&id)` (:43-65)
    pub fn sample<T: Scalar + Default, I: Scalar>(&self, component: Component<T>, identifier: Component<I>, id: I) -> Result<T, Error> {
        let mut out = T::default();
        check(unsafe {
            sys::incerto_sim_sample(self.raw, component.handle, identifier.handle, &id as *const I as *const c_void,
                                    &mut out as *mut T as *mut c_void, std::mem::size_of::<T>())
        })?;
        Ok(out)
    }
    /// `sample_single::<C>()` (:79-93)
    pub fn sample_single<T: Scalar + Default>(&self, component: Component<T>) -> Result<T, Error> {
        let mut out = T::default();
        check(unsafe { sys::incerto_sim_sample_single(self.raw, component.handle, &mut out as *mut T as *mut c_void, std::mem::size_of::<T>()) })?;
        Ok(out)
    }
    /// `get_aggregate_time_series::<C, Out>()` (:226-258): borrows the engine's host copy of the series.
    pub fn get_aggregate_time_series<'a, Out: Copy>(&'a self, aggregate: &Aggregate) -> Result<TimeSeries<'a, Out>, Error> {
        let (mut time, mut values) = (std::ptr::null(), std::ptr::null());
        let (mut len, mut record_size, mut interval) = (0u64, 0u64, 0u64);
        check(unsafe { sys::incerto_sim_get_aggregate_time_series(self.raw, &aggregate.raw(), &mut time, &mut values, &mut len, &mut record_size, &mut interval) })?;
        assert_eq!(record_size as usize, std::mem::size_of::<Out>(), "Out is not the aggregate's output type");
        // SAFETY: the engine keeps both arrays alive until the next run / reset / drop of this handle (&'a self)
        Ok(unsafe { TimeSeries { time: std::slice::from_raw_parts(time, len as usize),
                                 values: std::slice::from_raw_parts(values as *const Out, len as usize), sample_interval: interval } })
    }
    /// `get_time_series::<C, I>(&id)` (:186-216)
    pub fn get_time_series<'a, T: Scalar, I: Scalar>(&'a self, component: Component<T>, identifier: Component<I>, id: I) -> Result<TimeSeries<'a, T>, Error> {
        let (mut time, mut values) = (std::ptr::null(), std::ptr::null());
        let (mut len, mut record_size, mut interval) = (0u64, 0u64, 0u64);
        check(unsafe {
            sys::incerto_sim_get_time_series(self.raw, component.handle, identifier.handle, &id as *const I as *const c_void,
                                             &mut time, &mut values, &mut len, &mut record_size, &mut interval)
        })?;
        assert_eq!(record_size as usize, std::mem::size_of::<T>());
        Ok(unsafe { TimeSeries { time: std::slice::from_raw_parts(time, len as usize),
                                 values: std::slice::from_raw_parts(values as *const T, len as usize), sample_interval: interval } })
    }
    /// StepNumber (src/plugins/step_number.rs:10-23): 1 before the first step.
    pub fn step_number(&self) -> Result<usize, Error> {
        let mut n = 0u64;
        check(unsafe { sys::incerto_sim_step_number(self.raw, &mut n) })?;
        Ok(n as usize)
    }
}
impl Drop for Simulation {
    fn drop(&mut self) {
        if !self.raw.is_null() { unsafe { sys::incerto_sim_destroy(self.raw) } }
    }
}
