"""Short, profiler-friendly pass of the hot path: assemble a device-generated plate and run a
fixed number of PCG iterations.  Used under ncu (launch list and --set full capture).

    python profiles/prof_step.py [nx ny iters [spmv_format]]
"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from magnetite_b200 import _lib, meshgen  # noqa: E402

nx, ny, iters = (int(a) for a in (sys.argv[1:4] + ["4000", "2000", "20"][len(sys.argv[1:4]):]))
fmt = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib = _lib.load()
ctx = _lib.Context(0)
dm = C.c_void_p()
_lib.check(lib.mag_devmesh_plate(ctx.handle, nx, ny, 2.0, 3.0, C.byref(dm)), "plate")
view = _lib.MagMesh()
_lib.check(lib.mag_devmesh_view(dm, C.byref(view)), "view")
m = meshgen.EXAMPLE_MATERIAL
mat = _lib.MagMaterial(m.youngs_modulus, m.poisson_ratio, m.part_thickness)
opt = _lib.default_options(max_iter=iters, check_every=iters, spmv_format=fmt)
import torch  # noqa: E402  (device buffers for the result)
N, E = int(view.n_nodes), int(view.n_elems)
out = [torch.empty(N, dtype=torch.float64, device="cuda") for _ in range(4)] + [torch.empty(E, dtype=torch.float64, device="cuda")]
res = _lib.MagResult(*(t.data_ptr() for t in out), None, 1)
for rep in range(2):
    st = _lib.MagStats()
    rc = lib.mag_solve(ctx.handle, C.byref(view), C.byref(mat), C.byref(opt), C.byref(res), C.byref(st))
    assert rc in (0, _lib.MAG_ERR_NOT_CONVERGED), _lib.last_error()
    print({k: round(v, 3) if isinstance(v, float) else v for k, v in st.as_dict().items()})
lib.mag_devmesh_free(dm)
ctx.close()
