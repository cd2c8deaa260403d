//! Raw bindings: a literal transcription of `include/subsweep_b200.h` (ABI version 1).
//! tests/test_host_logic.py::test_rust_sys_crate_declares_every_symbol keeps the two in step.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const SSW_ABI_VERSION: i32 = 1;

pub const SSW_OK: c_int = 0;
pub const SSW_E_INVALID: c_int = -1;
pub const SSW_E_CUDA: c_int = -2;
pub const SSW_E_DEADLOCK: c_int = -3;
pub const SSW_E_NOMEM: c_int = -4;
pub const SSW_E_COMM: c_int = -5;

pub const SSW_FACE_LOCAL: u8 = 0;
pub const SSW_FACE_BOUNDARY: u8 = 1;
pub const SSW_FACE_LOCAL_PERIODIC: u8 = 2;

pub const SSW_FLAG_NO_SCHEDULE_CACHE: u32 = 1 << 0;
pub const SSW_FLAG_NO_COMPILED_PATH: u32 = 1 << 1;
pub const SSW_FLAG_NO_PATCH_PATH: u32 = 1 << 2;
pub const SSW_FLAG_SHARED_DEVICE: u32 = 1 << 3;

pub const SSW_COLL_REDUCE_SCATTER: c_int = 1;
pub const SSW_COLL_ALL_GATHER: c_int = 2;
pub const SSW_PEER_HANDLE_BYTES: usize = 64;

#[repr(C)]
pub struct ssw_params {
    pub n_dirs: i32,
    pub dirs_xyz: *const f64,
    pub n_levels: i32,
    pub max_timestep_s: f64,
    pub timestep_safety_factor: f64,
    pub chemistry_timestep_safety_factor: f64,
    pub significant_rate_threshold_per_s: f64,
    pub prevent_cooling: i32,
    pub scale_factor: f64,
    pub check_deadlock: i32,
    pub device_id: i32,
    pub rank: i32,
    pub world_size: i32,
    pub flags: u32,
}

#[repr(C)]
pub struct ssw_grid {
    pub n_cells: u64,
    pub face_offsets: *const u64,
    pub face_area: *const f64,
    pub face_normal: *const f64,
    pub face_neighbour: *const i32,
    pub face_kind: *const u8,
    pub cell_size: *const f64,
    pub cell_volume: *const f64,
}

#[repr(C)]
pub struct ssw_handle {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum ssw_field {
    XHII = 0,
    TEMPERATURE = 1,
    TIMESTEP = 2,
    PHOTON_RATE = 3,
    CHANGE_TIMESCALE = 4,
    PHOTOIONIZATION_RATE = 5,
    HEATING_RATE = 6,
    RECOMBINATION_RATE = 7,
    COLLISIONAL_IONIZATION_RATE = 8,
    PREVIOUS_RATE = 9,
    DENSITY = 10,
    SOURCE = 11,
    IONIZATION_TIME = 12,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum ssw_stat {
    TASKS_SOLVED = 0,
    SINGLE_SWEEPS = 1,
    CHEM_CELLS = 2,
    CHEM_FAILURES = 3,
    SCHEDULE_BUILDS = 4,
    SCHEDULE_REPLAYS = 5,
    KERNEL_LAUNCHES = 6,
    WAVEFRONT_LEVELS = 7,
    CHEM_ATTEMPTS = 8,
    CHEM_MAX_DEPTH = 9,
    PATCH_MACRO_TILES = 10,
    PATCH_LEVELS = 11,
    PATCH_PHASES = 12,
    WALK_WINDOW = 13,
    WALK_NEAR_PERMILLE = 14,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ssw_time_series {
    pub hydrogen_ionization_mass_average: f64,
    pub hydrogen_ionization_volume_average: f64,
    pub temperature_mass_average: f64,
    pub temperature_volume_average: f64,
    pub photoionization_rate_volume_average: f64,
    pub weighted_photoionization_rate_volume_average: f64,
    pub total_mass: f64,
    pub total_volume: f64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ssw_timings {
    pub sweep_ms: f64,
    pub chemistry_ms: f64,
    pub update_levels_ms: f64,
    pub schedule_ms: f64,
    pub allreduce_ms: f64,
    pub sweep_kernel_ms: f64,
    pub sweep_kernel_launches: u64,
    pub sweep_kernel_tasks: u64,
    pub sweep_level_ms: [f64; 32],
    pub step_ms: f64,
    pub steps: u64,
    pub kernel_level_ms: [f64; 32],
    pub kernel_level_tasks: [u64; 32],
    pub kernel_level_launches: [u64; 32],
}

pub type ssw_allreduce_fn =
    Option<unsafe extern "C" fn(ctx: *mut c_void, buf: *mut f64, n: u64, cuda_stream: *mut c_void) -> c_int>;
pub type ssw_collective_fn = Option<
    unsafe extern "C" fn(ctx: *mut c_void, op: c_int, buf: *mut f64, n_per_rank: u64, cuda_stream: *mut c_void) -> c_int,
>;

extern "C" {
    pub fn ssw_create(
        params: *const ssw_params,
        grid: *const ssw_grid,
        density: *const f64,
        xhii: *const f64,
        temperature: *const f64,
        source: *const f64,
        out: *mut *mut ssw_handle,
    ) -> c_int;
    pub fn ssw_destroy(h: *mut ssw_handle);
    pub fn ssw_set_allreduce(h: *mut ssw_handle, f: ssw_allreduce_fn, ctx: *mut c_void) -> c_int;
    pub fn ssw_set_collectives(h: *mut ssw_handle, f: ssw_collective_fn, ctx: *mut c_void) -> c_int;
    pub fn ssw_set_cell_positions(h: *mut ssw_handle, xyz: *const f64) -> c_int;
    pub fn ssw_patch_note(h: *mut ssw_handle) -> *const c_char;
    pub fn ssw_set_directions(h: *mut ssw_handle, dirs_xyz: *const f64) -> c_int;

    pub fn ssw_peer_arena(h: *mut ssw_handle, base: *mut *mut c_void, bytes: *mut u64) -> c_int;
    pub fn ssw_peer_export(h: *mut ssw_handle, ipc_handle_out: *mut c_void) -> c_int;
    pub fn ssw_peer_attach_ipc(h: *mut ssw_handle, ipc_handles: *const c_void) -> c_int;
    pub fn ssw_peer_attach(h: *mut ssw_handle, arena_bases: *const *mut c_void) -> c_int;

    pub fn ssw_run_sweeps(h: *mut ssw_handle, time_elapsed_s: *mut f64) -> c_int;
    pub fn ssw_set_inputs(h: *mut ssw_handle, density: *const f64, source: *const f64) -> c_int;

    pub fn ssw_read(h: *mut ssw_handle, field: ssw_field, out: *mut f64) -> c_int;
    pub fn ssw_read_begin(h: *mut ssw_handle, field: ssw_field, out: *mut f64) -> c_int;
    pub fn ssw_sync(h: *mut ssw_handle) -> c_int;
    pub fn ssw_read_levels(h: *mut ssw_handle, out: *mut u8) -> c_int;
    pub fn ssw_read_chem_attempts(h: *mut ssw_handle, out: *mut u16) -> c_int;
    pub fn ssw_level_counts(h: *mut ssw_handle, out: *mut u64) -> c_int;
    pub fn ssw_lowest_allowed_level(h: *mut ssw_handle, out: *mut i32) -> c_int;
    pub fn ssw_time_series_compute(h: *mut ssw_handle, mass: *const f64, with_rates: i32, out: *mut ssw_time_series) -> c_int;

    pub fn ssw_single_sweep(h: *mut ssw_handle, level: i32) -> c_int;
    pub fn ssw_set_levels(h: *mut ssw_handle, levels: *const u8) -> c_int;
    pub fn ssw_set_change_timescale(h: *mut ssw_handle, tau: *const f64) -> c_int;
    pub fn ssw_update_timestep_levels(h: *mut ssw_handle) -> c_int;
    pub fn ssw_read_dir_state(h: *mut ssw_handle, which: i32, out: *mut f64) -> c_int;
    pub fn ssw_read_wavefront_levels(h: *mut ssw_handle, level: i32, dir: i32, out: *mut i32) -> c_int;
    pub fn ssw_get_stat(h: *mut ssw_handle, which: ssw_stat, out: *mut u64) -> c_int;
    pub fn ssw_get_timings(h: *mut ssw_handle, out: *mut ssw_timings) -> c_int;
    pub fn ssw_reset_timings(h: *mut ssw_handle) -> c_int;
    pub fn ssw_set_timing_level(h: *mut ssw_handle, level: i32) -> c_int;

    pub fn ssw_direction_shard(n_dirs: i32, world_size: i32, rank: i32, begin: *mut i32, end: *mut i32) -> c_int;
    pub fn ssw_patch_lattice(xyz: *const f64, n_cells: u64, target_cells: i32, patch_of: *mut u32) -> i32;
    pub fn ssw_direction_groups(dirs_xyz: *const f64, n_dirs: i32, max_per_group: i32, group_of: *mut i32) -> i32;
    pub fn ssw_patch_levels(upwind: *const u32, n_groups: i32, n_patches: i32, level_out: *mut u32) -> i32;
    pub fn ssw_level_from_timesteps(max_num_levels: i32, max_timestep: f64, desired: f64) -> i32;
    pub fn ssw_levels_in_sweep_order(max_num_levels: i32, lowest_allowed: i32, out: *mut i32, cap: i32) -> i32;
    pub fn ssw_chemistry_batch(
        device_id: i32,
        n: u64,
        xhii: *mut f64,
        temperature: *mut f64,
        density: *const f64,
        volume: *const f64,
        length: *const f64,
        rate: *const f64,
        timestep: *const f64,
        scale_factor: f64,
        safety_factor: f64,
        prevent_cooling: i32,
        timescale_out: *mut f64,
        process_out: *mut i32,
        depth_out: *mut i32,
        attempts_out: *mut u64,
    ) -> c_int;

    pub fn ssw_last_error() -> *const c_char;
    pub fn ssw_abi_version() -> i32;
}
