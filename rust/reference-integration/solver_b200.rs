//! solver_b200.rs — drop-in for the reference's `solver::run` (src/solver.rs:543-586) and
//! `post_processor::csv_output` (src/post_processor.rs:18-83) on top of libmagnetite_b200.so.
//!
//! Vendored into the reference's `src/` (the reference is a binary crate, so only a module inside it can name
//! `crate::datatypes::*` and `crate::error::MagnetiteError`); `rust/reference-integration/apply.sh` copies it there,
//! adds the `magnetite-b200-sys` dependency and switches the two call sites of `src/main.rs`.
//!
//! UNVERIFIED: there is no Rust toolchain in the build image; this file has never been compiled.  The Python and C++
//! host layers (magnetite_b200/solver.py, host/magnetite_host.cpp) do the same flatten → `mag_solve` → write-back
//! around the same ABI and are what the GPU tests drive.
#![allow(dead_code)]

use std::ffi::{CStr, CString};

use magnetite_b200_sys as sys;

use crate::datatypes::{Element, ModelMetadata, Node};
use crate::error::MagnetiteError;

pub const DOF: usize = sys::MAG_DOF;                        // solver.rs:17
pub const MAX_CG_ITER: u64 = sys::MAG_MAX_CG_ITER;          // solver.rs:18
pub const TARGET_CG_COST: f64 = sys::MAG_TARGET_CG_COST;    // solver.rs:19

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::mag_last_error()).to_string_lossy().into_owned() }
}

fn host_last_error() -> String {
    unsafe { CStr::from_ptr(sys::mag_host_last_error()).to_string_lossy().into_owned() }
}

/// `Vec<Node>` / `Vec<Element>` (datatypes.rs:8-20) as the SoA arrays of `mag_mesh`; `known` bit0 ux, 1 uy, 2 fx,
/// 3 fy ⇔ `Option::is_some`.
struct Flat {
    x: Vec<f64>,
    y: Vec<f64>,
    ux: Vec<f64>,
    uy: Vec<f64>,
    fx: Vec<f64>,
    fy: Vec<f64>,
    known: Vec<u8>,
    n0: Vec<u32>,
    n1: Vec<u32>,
    n2: Vec<u32>,
}

impl Flat {
    fn zeroed(n: usize, e: usize) -> Flat {
        Flat {
            x: vec![0.0; n],
            y: vec![0.0; n],
            ux: vec![0.0; n],
            uy: vec![0.0; n],
            fx: vec![0.0; n],
            fy: vec![0.0; n],
            known: vec![0u8; n],
            n0: Vec::with_capacity(e),
            n1: Vec::with_capacity(e),
            n2: Vec::with_capacity(e),
        }
    }
}

fn flatten(nodes: &[Node], elements: &[Element]) -> Result<Flat, MagnetiteError> {
    if nodes.len() > u32::MAX as usize {
        return Err(MagnetiteError::Solver("more than 2^32 - 1 nodes".to_string()));
    }
    let mut f = Flat::zeroed(nodes.len(), elements.len());
    for (i, nd) in nodes.iter().enumerate() {
        f.x[i] = nd.vertex.x;
        f.y[i] = nd.vertex.y;
        if let Some(v) = nd.ux {
            f.ux[i] = v;
            f.known[i] |= sys::MAG_KNOWN_UX;
        }
        if let Some(v) = nd.uy {
            f.uy[i] = v;
            f.known[i] |= sys::MAG_KNOWN_UY;
        }
        if let Some(v) = nd.fx {
            f.fx[i] = v;
            f.known[i] |= sys::MAG_KNOWN_FX;
        }
        if let Some(v) = nd.fy {
            f.fy[i] = v;
            f.known[i] |= sys::MAG_KNOWN_FY;
        }
    }
    for el in elements {
        for &k in el.nodes.iter() {
            if k >= nodes.len() {
                // the reference panics on the out-of-bounds index (solver.rs:189-191)
                return Err(MagnetiteError::Solver(format!("element references node {} of {}", k, nodes.len())));
            }
        }
        f.n0.push(el.nodes[0] as u32);
        f.n1.push(el.nodes[1] as u32);
        f.n2.push(el.nodes[2] as u32);
    }
    Ok(f)
}

/// The same mesh with node i renamed `new_of_old[i]`; element order and orientation untouched.
fn permuted(f: &Flat, new_of_old: &[u32]) -> Flat {
    let n = f.x.len();
    let mut g = Flat::zeroed(n, f.n0.len());
    for i in 0..n {
        let j = new_of_old[i] as usize;
        g.x[j] = f.x[i];
        g.y[j] = f.y[i];
        g.ux[j] = f.ux[i];
        g.uy[j] = f.uy[i];
        g.fx[j] = f.fx[i];
        g.fy[j] = f.fy[i];
        g.known[j] = f.known[i];
    }
    for e in 0..f.n0.len() {
        g.n0.push(new_of_old[f.n0[e] as usize]);
        g.n1.push(new_of_old[f.n1[e] as usize]);
        g.n2.push(new_of_old[f.n2[e] as usize]);
    }
    g
}

/// Drop-in for `solver::run` (src/solver.rs:543-586).  `MAGNETITE_B200_REORDER=1` in the environment renumbers the
/// nodes (reverse Cuthill-McKee) around the solve — for meshes in gmsh order; the caller's vectors keep their order.
pub fn run(
    nodes: &mut Vec<Node>,
    elements: &mut Vec<Element>,
    model_metadata: &ModelMetadata,
) -> Result<(), MagnetiteError> {
    let reorder = std::env::var("MAGNETITE_B200_REORDER").map(|v| v == "1").unwrap_or(false);
    run_with(nodes, elements, model_metadata, reorder)
}

pub fn run_with(
    nodes: &mut Vec<Node>,
    elements: &mut Vec<Element>,
    model_metadata: &ModelMetadata,
    reorder: bool,
) -> Result<(), MagnetiteError> {
    println!("info: building element stiffness matrices..."); // solver.rs:551
    println!("info: building total stiffness matrix..."); // solver.rs:570
    let mut f = flatten(nodes, elements)?;
    let (n, e) = (nodes.len(), elements.len());
    let mut new_of_old: Vec<u32> = Vec::new(); // empty: solved in the caller's numbering
    if reorder {
        new_of_old = vec![0u32; n];
        let (mut before, mut after) = (0u64, 0u64);
        let rc = unsafe {
            sys::mag_reorder_rcm(
                n as u64,
                e as u64,
                f.n0.as_ptr(),
                f.n1.as_ptr(),
                f.n2.as_ptr(),
                new_of_old.as_mut_ptr(),
                &mut before,
                &mut after,
            )
        };
        if rc != 0 {
            return Err(MagnetiteError::Solver(host_last_error()));
        }
        if after < before {
            f = permuted(&f, &new_of_old);
        } else {
            new_of_old.clear();
        }
    }
    let mesh = sys::mag_mesh {
        n_nodes: n as u64,
        n_elems: e as u64,
        x: f.x.as_ptr(),
        y: f.y.as_ptr(),
        n0: f.n0.as_ptr(),
        n1: f.n1.as_ptr(),
        n2: f.n2.as_ptr(),
        ux: f.ux.as_ptr(),
        uy: f.uy.as_ptr(),
        fx: f.fx.as_ptr(),
        fy: f.fy.as_ptr(),
        known: f.known.as_ptr(),
        on_device: 0,
    };
    let mat = sys::mag_material {
        youngs_modulus: model_metadata.youngs_modulus,
        poisson_ratio: model_metadata.poisson_ratio,
        part_thickness: model_metadata.part_thickness,
    };
    let mut ux: Vec<f64> = vec![0.0; n];
    let mut uy: Vec<f64> = vec![0.0; n];
    let mut fx: Vec<f64> = vec![0.0; n];
    let mut fy: Vec<f64> = vec![0.0; n];
    let mut stress: Vec<f64> = vec![0.0; e];
    let mut out = sys::mag_result {
        ux: ux.as_mut_ptr(),
        uy: uy.as_mut_ptr(),
        fx: fx.as_mut_ptr(),
        fy: fy.as_mut_ptr(),
        stress: stress.as_mut_ptr(),
        sigma: std::ptr::null_mut(),
        on_device: 0,
    };
    let mut stats = sys::mag_stats::default();

    let mut ctx: *mut sys::mag_ctx = std::ptr::null_mut();
    if unsafe { sys::mag_ctx_create(&mut ctx, 0) } != 0 {
        return Err(MagnetiteError::Solver(last_error()));
    }
    println!("info: solving..."); // solver.rs:437
    let (rc, message) = unsafe {
        let mut opt: sys::mag_options = std::mem::zeroed();
        sys::mag_options_default(&mut opt);
        opt.compat = 1; // reference semantics: plain CG, x0 = 0, absolute cost 1e-4 (solver.rs:143, 153-154)
        let rc = sys::mag_solve(ctx, &mesh, &mat, &opt, &mut out, &mut stats);
        let message = if rc != 0 { last_error() } else { String::new() };
        sys::mag_ctx_destroy(ctx);
        (rc, message)
    };
    if rc != 0 {
        // solver.rs:160-164
        return Err(MagnetiteError::Solver(format!("Conjugate Gradient error: {}", message)));
    }
    println!("info: finished conjugate gradient approximation in {} iterations", stats.iters); // solver.rs:101-104
    println!("info: solved system in {:.3} seconds", f64::from(stats.ms_solve) / 1e3); // solver.rs:441
    for (i, node) in nodes.iter_mut().enumerate() {
        // solver.rs:476-482
        let j = if new_of_old.is_empty() { i } else { new_of_old[i] as usize };
        node.ux = Some(ux[j]);
        node.uy = Some(uy[j]);
        node.fx = Some(fx[j]);
        node.fy = Some(fy[j]);
    }
    for (i, el) in elements.iter_mut().enumerate() {
        // solver.rs:532-533
        el.stress = Some(stress[i]);
    }
    println!("info: solve complete"); // solver.rs:484
    Ok(())
}

/// `solver::compute_element_area` for every element at once (solver.rs:187-193; `mesher::check_ccw`,
/// mesher.rs:522-526, calls the scalar version per element and can keep doing so).
pub fn compute_element_areas(nodes: &Vec<Node>, elements: &Vec<Element>) -> Result<Vec<f64>, MagnetiteError> {
    let f = flatten(nodes, elements)?;
    let mesh = sys::mag_mesh {
        n_nodes: nodes.len() as u64,
        n_elems: elements.len() as u64,
        x: f.x.as_ptr(),
        y: f.y.as_ptr(),
        n0: f.n0.as_ptr(),
        n1: f.n1.as_ptr(),
        n2: f.n2.as_ptr(),
        ux: f.ux.as_ptr(),
        uy: f.uy.as_ptr(),
        fx: f.fx.as_ptr(),
        fy: f.fy.as_ptr(),
        known: f.known.as_ptr(),
        on_device: 0,
    };
    let mut area: Vec<f64> = vec![0.0; elements.len()];
    let mut ctx: *mut sys::mag_ctx = std::ptr::null_mut();
    if unsafe { sys::mag_ctx_create(&mut ctx, 0) } != 0 {
        return Err(MagnetiteError::Solver(last_error()));
    }
    let (rc, message) = unsafe {
        let rc = sys::mag_element_area(ctx, &mesh, area.as_mut_ptr());
        let message = if rc != 0 { last_error() } else { String::new() };
        sys::mag_ctx_destroy(ctx);
        (rc, message)
    };
    if rc != 0 {
        return Err(MagnetiteError::Solver(message));
    }
    Ok(area)
}

/// Drop-in for `post_processor::csv_output` (src/post_processor.rs:18-83): same files byte for byte (Rust `{}` float
/// formatting is reproduced in the library), written through 1 MiB buffers instead of one unbuffered write per row.
pub fn csv_output(
    elements: &Vec<Element>,
    nodes: &Vec<Node>,
    nodes_output: &str,
    elements_output: &str,
) -> Result<(), MagnetiteError> {
    let x: Vec<f64> = nodes.iter().map(|n| n.vertex.x).collect();
    let y: Vec<f64> = nodes.iter().map(|n| n.vertex.y).collect();
    let ux: Vec<f64> = nodes.iter().map(|n| n.ux.unwrap()).collect(); // the reference unwraps too (:50-51)
    let uy: Vec<f64> = nodes.iter().map(|n| n.uy.unwrap()).collect();
    let n0: Vec<u32> = elements.iter().map(|e| e.nodes[0] as u32).collect();
    let n1: Vec<u32> = elements.iter().map(|e| e.nodes[1] as u32).collect();
    let n2: Vec<u32> = elements.iter().map(|e| e.nodes[2] as u32).collect();
    let stress: Vec<f64> = elements.iter().map(|e| e.stress.unwrap()).collect(); // :70
    let nul = |_: std::ffi::NulError| MagnetiteError::Solver("output path contains a NUL byte".to_string());
    let np = CString::new(nodes_output).map_err(nul)?;
    let ep = CString::new(elements_output).map_err(nul)?;
    let rc = unsafe {
        sys::mag_csv_output(
            np.as_ptr(),
            ep.as_ptr(),
            x.len() as u64,
            x.as_ptr(),
            y.as_ptr(),
            ux.as_ptr(),
            uy.as_ptr(),
            n0.len() as u64,
            n0.as_ptr(),
            n1.as_ptr(),
            n2.as_ptr(),
            stress.as_ptr(),
        )
    };
    if rc != 0 {
        return Err(MagnetiteError::Solver(host_last_error())); // post_processor.rs:26-38 wraps these in Solver too
    }
    println!("info: wrote output to {} and {}", nodes_output, elements_output); // :77-80
    Ok(())
}
