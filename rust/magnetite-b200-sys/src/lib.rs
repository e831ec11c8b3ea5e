//! `extern "C"` binding of libmagnetite_b200.so: one `#[repr(C)]` struct per struct and one prototype per
//! function of include/magnetite_b200.h (ABI version 3), nothing else.  No dependencies.
//!
//! The safe wrapper with the signature of the reference's `solver::run` (src/solver.rs:543-547) is
//! `rust/reference-integration/solver_b200.rs`, a module a maintainer vendors into the reference's `src/` so that it
//! can name `crate::datatypes::*` and `crate::error::MagnetiteError` (the reference is a binary crate: an external
//! crate cannot see its types).
//!
//! UNVERIFIED: there is no Rust toolchain in the build image; this file has never been compiled.  What IS checked
//! (tests/test_rust_binding.py, on CPU): every struct field and every prototype below against the header — names,
//! order and types — and the struct sizes asserted in `mod layout` against the ctypes binding, which the GPU tests
//! drive through the same ABI.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const MAG_ABI_VERSION: c_int = 3;

// src/solver.rs:17-19
pub const MAG_DOF: usize = 2;
pub const MAG_MAX_CG_ITER: u64 = 10_000_000;
pub const MAG_TARGET_CG_COST: f64 = 1e-4;

pub const MAG_OK: c_int = 0;
pub const MAG_ERR_CUDA: c_int = -1;
pub const MAG_ERR_OOM: c_int = -2;
pub const MAG_ERR_BAD_BC: c_int = -3;
pub const MAG_ERR_BAD_INDEX: c_int = -4;
pub const MAG_ERR_INDEFINITE: c_int = -5;
pub const MAG_ERR_NOT_CONVERGED: c_int = -6;
pub const MAG_ERR_NCCL: c_int = -7;
pub const MAG_ERR_BAD_ARG: c_int = -8;

// which Option<f64> fields of Node (src/datatypes.rs:8-14) are Some(..)
pub const MAG_KNOWN_UX: u8 = 1;
pub const MAG_KNOWN_UY: u8 = 2;
pub const MAG_KNOWN_FX: u8 = 4;
pub const MAG_KNOWN_FY: u8 = 8;

#[repr(C)]
pub struct mag_mesh {
    pub n_nodes: u64,
    pub n_elems: u64,
    pub x: *const f64,
    pub y: *const f64,
    pub n0: *const u32,
    pub n1: *const u32,
    pub n2: *const u32,
    pub ux: *const f64,
    pub uy: *const f64,
    pub fx: *const f64,
    pub fy: *const f64,
    pub known: *const u8,
    pub on_device: i32,
}

#[repr(C)]
pub struct mag_material {
    pub youngs_modulus: f64,
    pub poisson_ratio: f64,
    pub part_thickness: f64,
}

#[repr(C)]
pub struct mag_options {
    pub rel_tol: f64,
    pub abs_tol: f64,
    pub max_iter: u64,
    pub precond: i32,
    pub compat: i32,
    pub cost_kind: i32,
    pub drop_exact_zeros: i32,
    pub check_every: i32,
    pub spmv_format: i32,
    pub want_sigma: i32,
    pub allreduce: i32,
    pub coarse_aggregates: i32,
    pub assembly: i32,
    pub result_scope: i32,
    pub reserved0: i32,
    pub stream: *mut c_void,
}

#[repr(C)]
pub struct mag_result {
    pub ux: *mut f64,
    pub uy: *mut f64,
    pub fx: *mut f64,
    pub fy: *mut f64,
    pub stress: *mut f64,
    pub sigma: *mut f64,
    pub on_device: i32,
}

#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct mag_stats {
    pub n_nodes: u64,
    pub n_elems: u64,
    pub n_dof: u64,
    pub n_free: u64,
    pub n_constrained: u64,
    pub nnz_structural: u64,
    pub nnz: u64,
    pub sell_entries: u64,
    pub iters: u64,
    pub final_residual: f64,
    pub b_norm: f64,
    pub converged: i32,
    pub negative_definite: i32,
    pub ms_upload: f32,
    pub ms_elem: f32,
    pub ms_sort: f32,
    pub ms_reduce: f32,
    pub ms_bc: f32,
    pub ms_format: f32,
    pub ms_solve: f32,
    pub ms_post: f32,
    pub ms_download: f32,
    pub ms_total: f32,
    pub kernel_launches: u64,
    pub spmv_bytes: u64,
    pub prof: [f64; 8],
    pub ms_coarse_setup: f32,
    pub n_coarse: u32,
    pub sell_index_bits: u32,
    pub precond_used: u32,
}

// opaque handles
#[repr(C)]
pub struct mag_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct mag_system {
    _private: [u8; 0],
}
#[repr(C)]
pub struct mag_devmesh {
    _private: [u8; 0],
}

extern "C" {
    // ---- lifecycle
    pub fn mag_abi_version() -> c_int;
    pub fn mag_last_error() -> *const c_char;
    pub fn mag_device_count(count: *mut c_int) -> c_int;
    pub fn mag_ctx_create(ctx: *mut *mut mag_ctx, device: c_int) -> c_int;
    pub fn mag_ctx_destroy(ctx: *mut mag_ctx);
    pub fn mag_options_default(opt: *mut mag_options);

    // ---- the drop-in call: replaces solver::run (src/solver.rs:543-586)
    pub fn mag_solve(ctx: *mut mag_ctx, mesh: *const mag_mesh, mat: *const mag_material, opt: *const mag_options, out: *mut mag_result, stats: *mut mag_stats) -> c_int;

    // ---- phase-level entry points
    pub fn mag_assemble(ctx: *mut mag_ctx, mesh: *const mag_mesh, mat: *const mag_material, opt: *const mag_options, sys: *mut *mut mag_system, stats: *mut mag_stats) -> c_int;
    pub fn mag_system_solve(sys: *mut mag_system, opt: *const mag_options, out: *mut mag_result, stats: *mut mag_stats) -> c_int;
    pub fn mag_system_free(sys: *mut mag_system);
    pub fn mag_system_info(sys: *const mag_system, stats: *mut mag_stats) -> c_int;

    // ---- parity exports
    pub fn mag_system_export_kff(sys: *const mag_system, rowptr: *mut i64, col: *mut i32, val: *mut f64, rhs: *mut f64, free_map: *mut i64) -> c_int;
    pub fn mag_system_export_full(sys: *const mag_system, rowptr: *mut i64, col: *mut i32, val: *mut f64) -> c_int;
    pub fn mag_element_stiffness(ctx: *mut mag_ctx, mesh: *const mag_mesh, mat: *const mag_material, ke: *mut f64) -> c_int;
    pub fn mag_element_area(ctx: *mut mag_ctx, mesh: *const mag_mesh, area: *mut f64) -> c_int;
    pub fn mag_strain_displacement(ctx: *mut mag_ctx, mesh: *const mag_mesh, b: *mut f64) -> c_int;
    pub fn mag_stress_strain(poisson_ratio: f64, youngs_modulus: f64, d: *mut f64) -> c_int;
    pub fn mag_stress(ctx: *mut mag_ctx, mesh: *const mag_mesh, mat: *const mag_material, ux: *const f64, uy: *const f64, stress: *mut f64, sigma: *mut f64) -> c_int;
    pub fn mag_system_spmv(sys: *mut mag_system, format: c_int, x: *const f64, y: *mut f64) -> c_int;
    pub fn mag_system_spmv_bench(sys: *mut mag_system, format: c_int, reps: c_int, ms_per_spmv: *mut f32, algorithmic_bytes: *mut u64) -> c_int;
    pub fn mag_system_residual(sys: *mut mag_system, ux: *const f64, uy: *const f64, on_device: c_int, rr_owned: *mut f64, bb_owned: *mut f64) -> c_int;

    // ---- output stage: post_processor::csv_output (src/post_processor.rs:18-83), host only
    pub fn mag_csv_output(nodes_path: *const c_char, elements_path: *const c_char, n_nodes: u64, x: *const f64, y: *const f64, ux: *const f64, uy: *const f64, n_elems: u64, n0: *const u32, n1: *const u32, n2: *const u32, stress: *const f64) -> c_int;
    pub fn mag_format_f64(v: f64, out: *mut c_char) -> usize;
    pub fn mag_host_last_error() -> *const c_char;

    // ---- node renumbering, host only
    pub fn mag_reorder_rcm(n_nodes: u64, n_elems: u64, n0: *const u32, n1: *const u32, n2: *const u32, new_of_old: *mut u32, band_before: *mut u64, band_after: *mut u64) -> c_int;
    pub fn mag_mesh_band(n_nodes: u64, n_elems: u64, n0: *const u32, n1: *const u32, n2: *const u32, band: *mut u64) -> c_int;

    // ---- synthetic meshes generated on the device
    pub fn mag_devmesh_plate(ctx: *mut mag_ctx, nx: u32, ny: u32, h: f64, ux_right: f64, out: *mut *mut mag_devmesh) -> c_int;
    pub fn mag_devmesh_perforated(ctx: *mut mag_ctx, nx: u32, ny: u32, h: f64, pitch: u32, radius: u32, ux_right: f64, out: *mut *mut mag_devmesh) -> c_int;
    pub fn mag_devmesh_download(dm: *const mag_devmesh, x: *mut f64, y: *mut f64, n0: *mut u32, n1: *mut u32, n2: *mut u32, ux: *mut f64, uy: *mut f64, fx: *mut f64, fy: *mut f64, known: *mut u8) -> c_int;
    pub fn mag_devmesh_view(dm: *const mag_devmesh, view: *mut mag_mesh) -> c_int;
    pub fn mag_devmesh_free(dm: *mut mag_devmesh);

    // ---- multi-GPU: one process per GPU, contiguous row blocks
    pub fn mag_comm_unique_id(id128: *mut c_void) -> c_int;
    pub fn mag_comm_init(ctx: *mut mag_ctx, rank: c_int, nranks: c_int, id128: *const c_void) -> c_int;
    pub fn mag_comm_rank(ctx: *const mag_ctx, rank: *mut c_int, nranks: *mut c_int) -> c_int;
    pub fn mag_partition_nodes(n_nodes: u64, nranks: c_int, rank: c_int, lo: *mut u64, hi: *mut u64) -> c_int;
    pub fn mag_halo_plan(nranks: c_int, rank: c_int, row_lo: *const u32, ext_lo: *const u32, ext_hi: *const u32, seg_lo: *mut u32, seg_hi: *mut u32, seg_dst: *mut i32, capacity: i32, n_segs: *mut i32) -> c_int;

    // ---- debug entry points used by the GPU unit tests
    pub fn mag_debug_sort_pairs(ctx: *mut mag_ctx, keys: *mut u64, payload: *mut u32, n: u64, key_bits: c_int) -> c_int;
    pub fn mag_debug_exclusive_scan(ctx: *mut mag_ctx, input: *const u32, out: *mut u32, n: u64) -> c_int;
    pub fn mag_debug_virtual_solve(ctx: *mut mag_ctx, mesh: *const mag_mesh, mat: *const mag_material, opt: *const mag_options, nranks: c_int, out: *mut mag_result, stats: *mut mag_stats) -> c_int;
}

/// Struct sizes on LP64 (the values tests/test_rust_binding.py reads and compares with `ctypes.sizeof` of the Python
/// binding); `cargo test` checks them against what rustc lays out.
pub mod layout {
    pub const SIZEOF_MAG_MESH: usize = 104;
    pub const SIZEOF_MAG_MATERIAL: usize = 24;
    pub const SIZEOF_MAG_OPTIONS: usize = 80;
    pub const SIZEOF_MAG_RESULT: usize = 56;
    pub const SIZEOF_MAG_STATS: usize = 232;
}

#[cfg(test)]
mod tests {
    use super::*;
    use std::mem::size_of;

    #[test]
    fn struct_sizes_match_the_header() {
        assert_eq!(size_of::<mag_mesh>(), layout::SIZEOF_MAG_MESH);
        assert_eq!(size_of::<mag_material>(), layout::SIZEOF_MAG_MATERIAL);
        assert_eq!(size_of::<mag_options>(), layout::SIZEOF_MAG_OPTIONS);
        assert_eq!(size_of::<mag_result>(), layout::SIZEOF_MAG_RESULT);
        assert_eq!(size_of::<mag_stats>(), layout::SIZEOF_MAG_STATS);
    }

    #[test]
    fn library_speaks_this_abi() {
        assert_eq!(unsafe { mag_abi_version() }, MAG_ABI_VERSION);
    }
}
