//! Thin `extern "C"` binding of libmagnetite_b200.so plus a safe wrapper with the exact
//! signature of the reference's `solver::run` (src/solver.rs:543-547), so `main.rs:64`
//! only changes its `use`.  Mirrors include/magnetite_b200.h (ABI version 3).
//!
//! UNVERIFIED: there is no Rust toolchain in the build image; this file has never been
//! compiled.  The Python ctypes binding (magnetite_b200/_lib.py) exercises the same ABI.
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub mod sys {
    use super::*;

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
    #[derive(Default)]
    pub struct mag_stats {
        pub n_nodes: u64, pub n_elems: u64, pub n_dof: u64, pub n_free: u64, pub n_constrained: u64,
        pub nnz_structural: u64, pub nnz: u64, pub sell_entries: u64, pub iters: u64,
        pub final_residual: f64, pub b_norm: f64,
        pub converged: i32, pub negative_definite: i32,
        pub ms_upload: f32, pub ms_elem: f32, pub ms_sort: f32, pub ms_reduce: f32, pub ms_bc: f32,
        pub ms_format: f32, pub ms_solve: f32, pub ms_post: f32, pub ms_download: f32, pub ms_total: f32,
        pub kernel_launches: u64, pub spmv_bytes: u64,
        pub prof: [f64; 8],
        pub ms_coarse_setup: f32, pub n_coarse: u32, pub sell_index_bits: u32, pub precond_used: u32,
    }
    pub enum mag_ctx {}

    extern "C" {
        pub fn mag_last_error() -> *const c_char;
        pub fn mag_ctx_create(ctx: *mut *mut mag_ctx, device: c_int) -> c_int;
        pub fn mag_ctx_destroy(ctx: *mut mag_ctx);
        pub fn mag_options_default(opt: *mut mag_options);
        pub fn mag_solve(ctx: *mut mag_ctx, mesh: *const mag_mesh, mat: *const mag_material,
                         opt: *const mag_options, out: *mut mag_result, stats: *mut mag_stats) -> c_int;
        pub fn mag_element_area(ctx: *mut mag_ctx, mesh: *const mag_mesh, area: *mut f64) -> c_int;
        // host-only entry points (no CUDA device needed); errors through mag_host_last_error
        pub fn mag_host_last_error() -> *const c_char;
        pub fn mag_reorder_rcm(n_nodes: u64, n_elems: u64, n0: *const u32, n1: *const u32, n2: *const u32,
                               new_of_old: *mut u32, band_before: *mut u64, band_after: *mut u64) -> c_int;
        pub fn mag_csv_output(nodes_path: *const c_char, elements_path: *const c_char, n_nodes: u64,
                              x: *const f64, y: *const f64, ux: *const f64, uy: *const f64, n_elems: u64,
                              n0: *const u32, n1: *const u32, n2: *const u32, stress: *const f64) -> c_int;
    }
}

/// The reference's own types (src/datatypes.rs, src/error.rs) are used unchanged.
use crate_types::{Element, MagnetiteError, ModelMetadata, Node};
pub mod crate_types {
    // In the reference tree these are `crate::datatypes::*` and `crate::error::MagnetiteError`.
    pub use magnetite_types::*;
}

pub const DOF: usize = 2;                       // solver.rs:17
pub const MAX_CG_ITER: u64 = 1e7 as u64;        // solver.rs:18
pub const TARGET_CG_COST: f64 = 1e-4;           // solver.rs:19

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::mag_last_error()).to_string_lossy().into_owned() }
}

fn host_last_error() -> String {
    unsafe { CStr::from_ptr(sys::mag_host_last_error()).to_string_lossy().into_owned() }
}

struct Flat {
    x: Vec<f64>, y: Vec<f64>, ux: Vec<f64>, uy: Vec<f64>, fx: Vec<f64>, fy: Vec<f64>, known: Vec<u8>,
    n0: Vec<u32>, n1: Vec<u32>, n2: Vec<u32>,
}

fn flatten(nodes: &Vec<Node>, elements: &Vec<Element>) -> Flat {
    let n = nodes.len();
    let mut f = Flat { x: Vec::with_capacity(n), y: Vec::with_capacity(n), ux: vec![0.0; n], uy: vec![0.0; n],
                       fx: vec![0.0; n], fy: vec![0.0; n], known: vec![0u8; n],
                       n0: Vec::with_capacity(elements.len()), n1: Vec::with_capacity(elements.len()),
                       n2: Vec::with_capacity(elements.len()) };
    for (i, nd) in nodes.iter().enumerate() {
        f.x.push(nd.vertex.x);
        f.y.push(nd.vertex.y);
        if let Some(v) = nd.ux { f.ux[i] = v; f.known[i] |= 1; }
        if let Some(v) = nd.uy { f.uy[i] = v; f.known[i] |= 2; }
        if let Some(v) = nd.fx { f.fx[i] = v; f.known[i] |= 4; }
        if let Some(v) = nd.fy { f.fy[i] = v; f.known[i] |= 8; }
    }
    for el in elements {
        f.n0.push(el.nodes[0] as u32);
        f.n1.push(el.nodes[1] as u32);
        f.n2.push(el.nodes[2] as u32);
    }
    f
}

/// The same mesh with node i renamed `new_of_old[i]`; element order and orientation untouched.
fn permuted(f: &Flat, new_of_old: &[u32]) -> Flat {
    let n = f.x.len();
    let mut g = Flat { x: vec![0.0; n], y: vec![0.0; n], ux: vec![0.0; n], uy: vec![0.0; n], fx: vec![0.0; n],
                       fy: vec![0.0; n], known: vec![0u8; n], n0: Vec::with_capacity(f.n0.len()),
                       n1: Vec::with_capacity(f.n0.len()), n2: Vec::with_capacity(f.n0.len()) };
    for i in 0..n {
        let j = new_of_old[i] as usize;
        g.x[j] = f.x[i]; g.y[j] = f.y[i];
        g.ux[j] = f.ux[i]; g.uy[j] = f.uy[i]; g.fx[j] = f.fx[i]; g.fy[j] = f.fy[i];
        g.known[j] = f.known[i];
    }
    for e in 0..f.n0.len() {
        g.n0.push(new_of_old[f.n0[e] as usize]);
        g.n1.push(new_of_old[f.n1[e] as usize]);
        g.n2.push(new_of_old[f.n2[e] as usize]);
    }
    g
}

/// Drop-in for `solver::run` (src/solver.rs:543-586).  `MAGNETITE_B200_REORDER=1` in the environment
/// renumbers the nodes (reverse Cuthill-McKee) around the solve — for meshes in gmsh order; the
/// caller's vectors keep their order.
pub fn run(nodes: &mut Vec<Node>, elements: &mut Vec<Element>, model_metadata: &ModelMetadata)
    -> Result<(), MagnetiteError>
{
    let reorder = std::env::var("MAGNETITE_B200_REORDER").map(|v| v == "1").unwrap_or(false);
    run_with(nodes, elements, model_metadata, reorder)
}

pub fn run_with(nodes: &mut Vec<Node>, elements: &mut Vec<Element>, model_metadata: &ModelMetadata,
                reorder: bool) -> Result<(), MagnetiteError>
{
    println!("info: building element stiffness matrices...");
    println!("info: building total stiffness matrix...");
    let mut f = flatten(nodes, elements);
    let (n, e) = (nodes.len(), elements.len());
    let mut new_of_old: Vec<u32> = Vec::new();           // empty: solved in the caller's numbering
    if reorder {
        new_of_old = vec![0u32; n];
        let (mut before, mut after) = (0u64, 0u64);
        let rc = unsafe {
            sys::mag_reorder_rcm(n as u64, e as u64, f.n0.as_ptr(), f.n1.as_ptr(), f.n2.as_ptr(),
                                 new_of_old.as_mut_ptr(), &mut before, &mut after)
        };
        if rc != 0 { return Err(MagnetiteError::Solver(host_last_error())); }
        if after < before { f = permuted(&f, &new_of_old); } else { new_of_old.clear(); }
    }
    let mesh = sys::mag_mesh {
        n_nodes: n as u64, n_elems: e as u64, x: f.x.as_ptr(), y: f.y.as_ptr(),
        n0: f.n0.as_ptr(), n1: f.n1.as_ptr(), n2: f.n2.as_ptr(),
        ux: f.ux.as_ptr(), uy: f.uy.as_ptr(), fx: f.fx.as_ptr(), fy: f.fy.as_ptr(),
        known: f.known.as_ptr(), on_device: 0,
    };
    let mat = sys::mag_material { youngs_modulus: model_metadata.youngs_modulus,
                                  poisson_ratio: model_metadata.poisson_ratio,
                                  part_thickness: model_metadata.part_thickness };
    let (mut ux, mut uy, mut fx, mut fy) = (vec![0.0; n], vec![0.0; n], vec![0.0; n], vec![0.0; n]);
    let mut stress = vec![0.0; e];
    let mut out = sys::mag_result { ux: ux.as_mut_ptr(), uy: uy.as_mut_ptr(), fx: fx.as_mut_ptr(),
                                    fy: fy.as_mut_ptr(), stress: stress.as_mut_ptr(),
                                    sigma: std::ptr::null_mut(), on_device: 0 };
    let mut stats = sys::mag_stats::default();
    let rc = unsafe {
        let mut ctx: *mut sys::mag_ctx = std::ptr::null_mut();
        let rc = sys::mag_ctx_create(&mut ctx, 0);
        if rc != 0 { return Err(MagnetiteError::Solver(last_error())); }
        let mut opt: sys::mag_options = std::mem::zeroed();
        sys::mag_options_default(&mut opt);
        opt.compat = 1;                       // reference semantics: plain CG, absolute cost 1e-4
        println!("info: solving...");
        let rc = sys::mag_solve(ctx, &mesh, &mat, &opt, &mut out, &mut stats);
        sys::mag_ctx_destroy(ctx);
        rc
    };
    if rc != 0 {
        return Err(MagnetiteError::Solver(format!("Conjugate Gradient error: {}", last_error())));
    }
    println!("info: finished conjugate gradient approximation in {} iterations", stats.iters);
    println!("info: solved system in {:.3} seconds", stats.ms_solve / 1e3);
    for (i, node) in nodes.iter_mut().enumerate() {          // solver.rs:476-482
        let j = if new_of_old.is_empty() { i } else { new_of_old[i] as usize };
        node.ux = Some(ux[j]); node.uy = Some(uy[j]);
        node.fx = Some(fx[j]); node.fy = Some(fy[j]);
    }
    for (i, el) in elements.iter_mut().enumerate() {          // solver.rs:532-533
        el.stress = Some(stress[i]);
    }
    println!("info: solve complete");
    Ok(())
}

/// Drop-in for `post_processor::csv_output` (src/post_processor.rs:18-83): same files byte for byte
/// (Rust `{}` float formatting is reproduced in the library), written through 1 MiB buffers instead of
/// one unbuffered write per row.
pub fn csv_output(elements: &Vec<Element>, nodes: &Vec<Node>, nodes_output: &str, elements_output: &str)
    -> Result<(), MagnetiteError>
{
    use std::ffi::CString;
    let x: Vec<f64> = nodes.iter().map(|n| n.vertex.x).collect();
    let y: Vec<f64> = nodes.iter().map(|n| n.vertex.y).collect();
    let ux: Vec<f64> = nodes.iter().map(|n| n.ux.unwrap()).collect();       // the reference unwraps too (:50-51)
    let uy: Vec<f64> = nodes.iter().map(|n| n.uy.unwrap()).collect();
    let n0: Vec<u32> = elements.iter().map(|e| e.nodes[0] as u32).collect();
    let n1: Vec<u32> = elements.iter().map(|e| e.nodes[1] as u32).collect();
    let n2: Vec<u32> = elements.iter().map(|e| e.nodes[2] as u32).collect();
    let stress: Vec<f64> = elements.iter().map(|e| e.stress.unwrap()).collect();   // :70
    let bad_path = |_| MagnetiteError::Solver("output path contains a NUL byte".to_string());
    let np = CString::new(nodes_output).map_err(bad_path)?;
    let ep = CString::new(elements_output).map_err(bad_path)?;
    let rc = unsafe {
        sys::mag_csv_output(np.as_ptr(), ep.as_ptr(), x.len() as u64, x.as_ptr(), y.as_ptr(), ux.as_ptr(),
                            uy.as_ptr(), n0.len() as u64, n0.as_ptr(), n1.as_ptr(), n2.as_ptr(), stress.as_ptr())
    };
    if rc != 0 { return Err(MagnetiteError::Solver(host_last_error())); }
    println!("info: wrote output to {} and {}", nodes_output, elements_output);   // :77-80
    Ok(())
}
