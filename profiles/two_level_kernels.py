"""One two-level solve of the device-generated plate (default 4000 x 2000 = 16 M DOF) with a bounded number of
iterations: what `ncu --metrics gpu__time_duration.sum` is pointed at to get the per-kernel durations of the
coarse setup and of one iteration.   python profiles/two_level_kernels.py [nx ny max_iter [virtual_ranks]]
With virtual_ranks > 1 the solve runs as that many row blocks inside this process (mag_debug_virtual_solve): the
per-rank kernel shapes of a multi-GPU run, one after the other on one GPU."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from magnetite_b200 import _lib, meshgen  # noqa: E402

nx, ny, iters = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4000, 2000, 12)
vranks = int(sys.argv[4]) if len(sys.argv) > 4 else 1
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
opt = _lib.default_options(precond=2, max_iter=iters, check_every=iters)
st = _lib.MagStats()
rc = (lib.mag_solve(ctx.handle, C.byref(view), C.byref(mat), C.byref(opt), C.byref(res), C.byref(st)) if vranks == 1 else
      lib.mag_debug_virtual_solve(ctx.handle, C.byref(view), C.byref(mat), C.byref(opt), vranks, C.byref(res), C.byref(st)))
print(f"rc {rc} (-6 = stopped at max_iter, as intended), iterations {st.iters}, coarse setup {st.ms_coarse_setup:.1f} ms, "
      f"solve {st.ms_solve:.1f} ms, timeline {[round(v / 1e3, 1) for v in list(st.prof)[:7]]}")
lib.mag_devmesh_free(dm)
ctx.close()
