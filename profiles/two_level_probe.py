"""Iterations and time-to-solve of Jacobi-PCG vs the two-level preconditioner on device-generated plates.

    python profiles/two_level_probe.py [nx ny [aggregates ...]]
"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from magnetite_b200 import _lib, meshgen  # noqa: E402

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 500)
aggs = [int(a) for a in sys.argv[3:]] or [0]
lib = _lib.load()
ctx = _lib.Context(0)
dm = C.c_void_p()
_lib.check(lib.mag_devmesh_plate(ctx.handle, nx, ny, 2.0, 3.0, C.byref(dm)), "plate")
view = _lib.MagMesh()
_lib.check(lib.mag_devmesh_view(dm, C.byref(view)), "view")
m = meshgen.EXAMPLE_MATERIAL
mat = _lib.MagMaterial(m.youngs_modulus, m.poisson_ratio, m.part_thickness)
N, E = int(view.n_nodes), int(view.n_elems)
out = [torch.empty(N, dtype=torch.float64, device="cuda") for _ in range(4)] + [torch.empty(E, dtype=torch.float64, device="cuda")]
res = _lib.MagResult(*(t.data_ptr() for t in out), None, 1)
sysh = C.c_void_p(); st = _lib.MagStats()
_lib.check(lib.mag_assemble(ctx.handle, C.byref(view), C.byref(mat), C.byref(_lib.default_options()), C.byref(sysh), C.byref(st)), "assemble")
ref = None
for label, opt in [("jacobi", _lib.default_options())] + [(f"two-level/{a or 'auto'}", _lib.default_options(precond=2, coarse_aggregates=a)) for a in aggs]:
    for rep in range(2):                       # the first two-level solve also builds the coarse space
        st = _lib.MagStats()
        _lib.check(lib.mag_system_solve(sysh, C.byref(opt), C.byref(res), C.byref(st)), "solve")
        ux = out[0].clone()
        if ref is None:
            ref = ux
        err = float((ux - ref).norm() / ref.norm())
        print(f"plate {nx}x{ny} {label:16s} rep {rep}: {st.iters:6d} it, solve {st.ms_solve:9.1f} ms, "
              f"{1e3 * st.ms_solve / max(st.iters, 1):7.1f} us/it, rel.res {st.final_residual / st.b_norm:.2e}, |du|/|u| vs jacobi {err:.1e}",
              flush=True)
lib.mag_system_free(sysh)
lib.mag_devmesh_free(dm)
ctx.close()
