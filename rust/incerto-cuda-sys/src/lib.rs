//! UNBUILT IN THIS ENVIRONMENT (no Rust toolchain): maintainer-side binding of include/incerto_b200.h.
//! Hand-written `extern "C"` block (bindgen is unavailable); one declaration per exported symbol.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub mod philox_shim;

#[repr(C)] pub struct incerto_builder { _p: [u8; 0] }
#[repr(C)] pub struct incerto_sim { _p: [u8; 0] }
pub type incerto_component = i32;
pub type incerto_grid = i32;

#[repr(C)] pub struct incerto_column_data { pub component: incerto_component, pub data: *const c_void }
#[repr(C)] pub struct incerto_system_desc { pub kind: u32, pub params: *const c_void, pub params_size: u32 }
#[repr(C)] pub struct incerto_filter { pub with: *const incerto_component, pub n_with: u32, pub without: *const incerto_component, pub n_without: u32 }
#[repr(C)] pub struct incerto_aggregate_desc { pub component: incerto_component, pub component2: incerto_component, pub kind: u32, pub percentile: u32, pub param: u64, pub filter: incerto_filter }
#[repr(C)] pub struct incerto_kernel_stat { pub name: *const c_char, pub launches: u64, pub device_ms: f64, pub algorithmic_bytes: f64 }

extern "C" {
    pub fn incerto_last_error_message() -> *const c_char;
    pub fn incerto_abi_version() -> u32;
    pub fn incerto_builder_new(out: *mut *mut incerto_builder) -> i32;
    pub fn incerto_builder_free(b: *mut incerto_builder);
    pub fn incerto_builder_set_seed(b: *mut incerto_builder, seed: u64, replica: u32) -> i32;
    pub fn incerto_builder_set_device(b: *mut incerto_builder, ordinal: i32) -> i32;
    pub fn incerto_builder_set_row_partition(b: *mut incerto_builder, rank: u32, world: u32) -> i32;
    pub fn incerto_builder_register_component(b: *mut incerto_builder, name: *const c_char, dtype: i32, lanes: u32, out: *mut incerto_component) -> i32;
    pub fn incerto_builder_add_entity_spawner(b: *mut incerto_builder, n: u64, columns: *const incerto_column_data, n_columns: u32) -> i32;
    pub fn incerto_builder_add_entity_spawner_borrowed(b: *mut incerto_builder, n: u64, columns: *const incerto_column_data, n_columns: u32) -> i32;
    pub fn incerto_builder_add_device_spawner(b: *mut incerto_builder, kind: u32, params: *const c_void, params_size: u32) -> i32;
    pub fn incerto_builder_add_systems(b: *mut incerto_builder, systems: *const incerto_system_desc, n: u32) -> i32;
    pub fn incerto_builder_add_spatial_grid(b: *mut incerto_builder, dim: u32, bounds: *const i32, position: incerto_component, component: incerto_component, out: *mut incerto_grid) -> i32;
    pub fn incerto_builder_record_aggregate_time_series(b: *mut incerto_builder, agg: *const incerto_aggregate_desc, sample_interval: u64) -> i32;
    pub fn incerto_builder_record_time_series(b: *mut incerto_builder, component: incerto_component, identifier: incerto_component, sample_interval: u64) -> i32;
    pub fn incerto_builder_build(b: *mut incerto_builder, out: *mut *mut incerto_sim) -> i32;
    pub fn incerto_sim_run(s: *mut incerto_sim, num_steps: u64) -> i32;
    pub fn incerto_sim_run_async(s: *mut incerto_sim, num_steps: u64) -> i32;
    pub fn incerto_sim_synchronize(s: *mut incerto_sim) -> i32;
    pub fn incerto_sim_destroy(s: *mut incerto_sim);
    pub fn incerto_sim_reset(s: *mut incerto_sim, replica: u32) -> i32;
    pub fn incerto_sim_count(s: *mut incerto_sim, filter: *const incerto_filter, out: *mut u64) -> i32;
    pub fn incerto_sim_sample_aggregate(s: *mut incerto_sim, agg: *const incerto_aggregate_desc, out: *mut c_void, out_size: usize) -> i32;
    pub fn incerto_sim_sample_aggregate_global(s: *mut incerto_sim, agg: *const incerto_aggregate_desc, out: *mut c_void, out_size: usize) -> i32;
    pub fn incerto_sim_sample(s: *mut incerto_sim, component: incerto_component, identifier: incerto_component, id_value: *const c_void, out: *mut c_void, out_size: usize) -> i32;
    pub fn incerto_sim_sample_single(s: *mut incerto_sim, component: incerto_component, out: *mut c_void, out_size: usize) -> i32;
    pub fn incerto_sim_get_aggregate_time_series(s: *mut incerto_sim, agg: *const incerto_aggregate_desc, time: *mut *const u64, values: *mut *const c_void, len: *mut u64, record_size: *mut u64, sample_interval: *mut u64) -> i32;
    pub fn incerto_sim_get_time_series(s: *mut incerto_sim, component: incerto_component, identifier: incerto_component, id_value: *const c_void, time: *mut *const u64, values: *mut *const c_void, len: *mut u64, record_size: *mut u64, sample_interval: *mut u64) -> i32;
    pub fn incerto_sim_grid_entities_at(s: *mut incerto_sim, g: incerto_grid, pos: *const i32, out_uids: *mut u64, capacity: u64, n_out: *mut u64) -> i32;
    pub fn incerto_sim_grid_neighbors_of(s: *mut incerto_sim, g: incerto_grid, pos: *const i32, orthogonal: u32, out_uids: *mut u64, capacity: u64, n_out: *mut u64) -> i32;
    pub fn incerto_sim_grid_is_empty(s: *mut incerto_sim, g: incerto_grid, pos: *const i32, out: *mut u32) -> i32;
    pub fn incerto_sim_grid_num_entities(s: *mut incerto_sim, g: incerto_grid, out: *mut u64) -> i32;
    pub fn incerto_sim_grid_position_of(s: *mut incerto_sim, g: incerto_grid, uid: u64, out_pos: *mut i32) -> i32;
    pub fn incerto_sim_grid_in_bounds(s: *mut incerto_sim, g: incerto_grid, pos: *const i32, out: *mut u32) -> i32;
    pub fn incerto_sim_step_number(s: *mut incerto_sim, out: *mut u64) -> i32;
    pub fn incerto_sim_num_entities(s: *mut incerto_sim, component: incerto_component, out: *mut u64) -> i32;
    pub fn incerto_sim_read_component(s: *mut incerto_sim, component: incerto_component, out: *mut c_void, out_size: usize, uids_out: *mut u64) -> i32;
    pub fn incerto_sim_read_marker(s: *mut incerto_sim, marker: incerto_component, rows_of: incerto_component, out: *mut u8, out_size: usize) -> i32;
    pub fn incerto_sim_last_run_stats(s: *mut incerto_sim, device_ms: *mut f64, kernel_launches: *mut u64) -> i32;
    pub fn incerto_sim_set_profiling(s: *mut incerto_sim, enabled: u32) -> i32;
    pub fn incerto_sim_kernel_stats(s: *mut incerto_sim, out: *mut incerto_kernel_stat, capacity: u32, n_out: *mut u32) -> i32;
    pub fn incerto_sim_flush_l2(s: *mut incerto_sim) -> i32;
    pub fn incerto_sim_run_cold(s: *mut incerto_sim, num_steps: u64) -> i32;
    pub fn incerto_sim_span_ms(first: *mut incerto_sim, last: *mut incerto_sim, ms_out: *mut f64) -> i32;
    pub fn incerto_comm_unique_id(out: *mut u8) -> i32;
    pub fn incerto_sim_attach_comm(s: *mut incerto_sim, unique_id: *const u8, rank: u32, world: u32) -> i32;
    pub fn incerto_sim_halo_export(s: *mut incerto_sim, side: u32, out: *mut c_void, capacity: usize, n_out: *mut usize) -> i32;
    pub fn incerto_sim_halo_import(s: *mut incerto_sim, side: u32, input: *const c_void, n: usize) -> i32;
    pub fn incerto_partition_range(width: i32, rank: u32, world: u32, begin: *mut i32, end: *mut i32) -> i32;
    pub fn incerto_sim_allreduce_u64(s: *mut incerto_sim, values: *mut u64, n: u32, op: u32) -> i32;
    pub fn incerto_sim_allreduce_f64(s: *mut incerto_sim, values: *mut f64, n: u32, op: u32) -> i32;
}

/// src/error.rs:12-43 rebuilt from the status code.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum SamplingError { ComponentDoesNotExist = 1, SingleNoEntities, SingleMultipleEntities, AggregateNoEntities, TimeSeriesNotRecorded, EntityIdentifierNotFound, EntityIdentifierNotUnique }
/// src/error.rs:47-52
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum BuilderError { TimeSeriesRecordingConflict = 16 }
