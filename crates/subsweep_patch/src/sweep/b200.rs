//! `src/sweep/b200.rs` -- drop this file into subsweep's tree next to `src/sweep/mod.rs` and apply the edits of
//! `crates/subsweep_patch/README.md`.  It replaces the `Sweep<HydrogenOnly>` resource (src/sweep/mod.rs:159-179) by a
//! handle to libsubsweep_b200.so; the plugin surface (SweepPlugin, the `sweep:` parameters, the per-particle
//! components and what `run_sweep_system` writes back) is unchanged.
//!
//! Not compiled in this repository (no Rust toolchain in the build image); `include/subsweep_b200.h` is the contract
//! and `tests/c_abi_smoke.c` the compiled proof that the header and the library agree.
use subsweep_b200_sys as ffi;

use super::direction::Directions;
use super::grid::Cell;
use super::grid::ParticleType;
use super::parameters::SweepParameters;
use crate::units::Time;
use crate::units::VecLength;

/// Owns the device-side solver.  Single caller, like the NonSend resource it replaces (mod.rs:139).
pub struct B200Sweep {
    h: *mut ffi::ssw_handle,
    n: usize,
    scratch: Vec<f64>,
}

fn check(rc: i32) {
    if rc != ffi::SSW_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::ssw_last_error()) }
            .to_string_lossy()
            .into_owned();
        // the reference panics as well (unwrap / assert!); SSW_E_DEADLOCK is where Sweep::solve would spin (mod.rs:291-300)
        panic!("libsubsweep_b200: {msg} ({rc})");
    }
}

fn flatten_directions(directions: &Directions) -> Vec<f64> {
    directions.iter().flat_map(|d| [d.x, d.y, d.z]).collect()
}

impl B200Sweep {
    /// `cells`, `positions` and the four field slices in `ParticleId.index` order 0..N-1 (what `ActiveList::new`
    /// requires of the reference's own containers, active_list.rs:25-35).  Remote / RemotePeriodic neighbours do not
    /// occur: every process holds the whole grid (`WorldSize` = 1 from bevy's point of view) and `rank` / `world_size`
    /// only select the direction shard.
    #[allow(clippy::too_many_arguments)]
    pub fn new(
        directions: &Directions,
        cells: &[Cell],
        positions: Option<&[VecLength]>,
        density: &[f64],
        xhii: &[f64],
        temperature: &[f64],
        source: &[f64],
        p: &SweepParameters,
        scale_factor: f64,
        rank: i32,
        world_size: i32,
    ) -> Self {
        let n = cells.len();
        assert!(density.len() == n && xhii.len() == n && temperature.len() == n && source.len() == n);
        let mut off = Vec::with_capacity(n + 1);
        off.push(0u64);
        let (mut area, mut normal, mut nb, mut kind) = (vec![], vec![], vec![], vec![]);
        let (mut size, mut volume) = (Vec::with_capacity(n), Vec::with_capacity(n));
        for cell in cells {
            for (face, ptype) in &cell.neighbours {
                // grid/cell.rs:92-97, 119-123; face order is kept: it is the summation order of solve_task
                area.push(face.area.value_unchecked());
                normal.extend_from_slice(&[face.normal.x, face.normal.y, face.normal.z]);
                let (k, id) = match ptype {
                    ParticleType::Local(id) => (ffi::SSW_FACE_LOCAL, id.index as i32),
                    ParticleType::LocalPeriodic(pn) => (ffi::SSW_FACE_LOCAL_PERIODIC, pn.id.index as i32),
                    ParticleType::Boundary => (ffi::SSW_FACE_BOUNDARY, -1),
                    _ => unreachable!("one process per GPU holds the whole grid: no remote neighbours"),
                };
                kind.push(k);
                nb.push(id);
            }
            off.push(area.len() as u64);
            size.push(cell.size.value_unchecked());
            volume.push(cell.volume.value_unchecked());
        }
        let dirs = flatten_directions(directions);
        let params = ffi::ssw_params {
            n_dirs: directions.len() as i32,
            dirs_xyz: dirs.as_ptr(),
            n_levels: p.num_timestep_levels as i32,
            max_timestep_s: p.max_timestep.value_unchecked(),
            timestep_safety_factor: p.timestep_safety_factor.value_unchecked(),
            chemistry_timestep_safety_factor: p.chemistry_timestep_safety_factor.value_unchecked(),
            significant_rate_threshold_per_s: p.significant_rate_threshold.value_unchecked(),
            prevent_cooling: p.prevent_cooling as i32,
            scale_factor,
            check_deadlock: p.check_deadlock as i32,
            device_id: rank, // one process per GPU of the box
            rank,
            world_size,
            flags: 0,
        };
        let grid = ffi::ssw_grid {
            n_cells: n as u64,
            face_offsets: off.as_ptr(),
            face_area: area.as_ptr(),
            face_normal: normal.as_ptr(),
            face_neighbour: nb.as_ptr(),
            face_kind: kind.as_ptr(),
            cell_size: size.as_ptr(),
            cell_volume: volume.as_ptr(),
        };
        let mut h = std::ptr::null_mut();
        check(unsafe {
            ffi::ssw_create(
                &params,
                &grid,
                density.as_ptr(),
                xhii.as_ptr(),
                temperature.as_ptr(),
                source.as_ptr(),
                &mut h,
            )
        });
        // Optional: the Position component lets the library order the all-cells sweep patch by patch (DESIGN.md
        // section 5.3).  The solver itself never needs positions.
        if let Some(positions) = positions {
            assert_eq!(positions.len(), n);
            let xyz: Vec<f64> = positions
                .iter()
                .flat_map(|p| [p.x().value_unchecked(), p.y().value_unchecked(), p.z().value_unchecked()])
                .collect();
            check(unsafe { ffi::ssw_set_cell_positions(h, xyz.as_ptr()) });
        }
        Self { h, n, scratch: vec![0.0; n] }
    }

    /// Direction sharding over the GPUs of one box without a collective library: every rank exports its arena
    /// (`ssw_peer_export`), the 64-byte handles are all-gathered over the host communicator the application already
    /// has (`MpiWorld::all_gather`, src/communication), and every rank maps the others' arenas.
    pub fn attach_peers(&mut self, all_gather_handles: impl Fn(&[u8; ffi::SSW_PEER_HANDLE_BYTES]) -> Vec<u8>) {
        let mut mine = [0u8; ffi::SSW_PEER_HANDLE_BYTES];
        check(unsafe { ffi::ssw_peer_export(self.h, mine.as_mut_ptr() as *mut _) });
        let all = all_gather_handles(&mine);
        check(unsafe { ffi::ssw_peer_attach_ipc(self.h, all.as_ptr() as *const _) });
    }

    /// Sweep::run_sweeps (mod.rs:258-272)
    pub fn run_sweeps(&mut self) -> Time {
        let mut t = 0.0;
        check(unsafe { ffi::ssw_run_sweeps(self.h, &mut t) });
        Time::seconds(t)
    }

    /// One component, blocking (N doubles in ParticleId.index order).
    pub fn read(&mut self, f: ffi::ssw_field) -> &[f64] {
        check(unsafe { ffi::ssw_read(self.h, f, self.scratch.as_mut_ptr()) });
        &self.scratch
    }

    /// The five components `run_sweep_system` writes back, queued behind each other and waited for once.
    pub fn read_all(&mut self, out: &mut [Vec<f64>; 5]) {
        const FIELDS: [ffi::ssw_field; 5] = [
            ffi::ssw_field::XHII,
            ffi::ssw_field::TEMPERATURE,
            ffi::ssw_field::TIMESTEP,
            ffi::ssw_field::PHOTON_RATE,
            ffi::ssw_field::IONIZATION_TIME,
        ];
        for (f, buf) in FIELDS.iter().zip(out.iter_mut()) {
            buf.resize(self.n, 0.0);
            check(unsafe { ffi::ssw_read_begin(self.h, *f, buf.as_mut_ptr()) });
        }
        check(unsafe { ffi::ssw_sync(self.h) });
    }

    /// `Source` / `Density` components changed on the host (new sources, remapped densities).
    pub fn set_inputs(&mut self, density: Option<&[f64]>, source: Option<&[f64]>) {
        let d = density.map_or(std::ptr::null(), |v| v.as_ptr());
        let s = source.map_or(std::ptr::null(), |v| v.as_ptr());
        check(unsafe { ffi::ssw_set_inputs(self.h, d, s) });
    }

    /// rotate_directions_system (direction/mod.rs:158-174): the host rotates the bins, the library remaps the fluxes.
    pub fn set_directions(&mut self, directions: &Directions) {
        let dirs = flatten_directions(directions);
        check(unsafe { ffi::ssw_set_directions(self.h, dirs.as_ptr()) });
    }

    /// num_particles_at_timestep_levels_system (mod.rs:765-780): cumulative counts per level.
    pub fn level_counts(&mut self, num_levels: usize) -> Vec<u64> {
        let mut out = vec![0u64; num_levels];
        check(unsafe { ffi::ssw_level_counts(self.h, out.as_mut_ptr()) });
        out
    }

    /// compute_time_series_system (src/sweep/time_series.rs): the averages reduced on the device.
    pub fn time_series(&mut self, mass: &[f64], with_rates: bool) -> ffi::ssw_time_series {
        assert_eq!(mass.len(), self.n);
        let mut out = ffi::ssw_time_series::default();
        check(unsafe { ffi::ssw_time_series_compute(self.h, mass.as_ptr(), with_rates as i32, &mut out) });
        out
    }
}

impl Drop for B200Sweep {
    fn drop(&mut self) {
        unsafe { ffi::ssw_destroy(self.h) }
    }
}
