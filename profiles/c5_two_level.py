"""torchrun worker: BASELINE configs[4] (perforated plate, 16 M DOF per GPU before the holes) through the
opt-in two-level preconditioner, next to the Jacobi number bench.py --workload c5 reports.

    python -m torch.distributed.run --nproc-per-node 8 profiles/c5_two_level.py
"""
import ctypes as C
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from magnetite_b200 import _lib, dist as mdist, meshgen  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    rank, world, local = mdist.init_process_group("nccl")
else:
    rank, local = 0, 0
torch.cuda.set_device(local)
lib = _lib.load()
ctx = _lib.Context(local)
if world > 1:
    mdist.init_comm(ctx)
nx, ny = int(round(4000 * world ** 0.5)), int(round(2000 * world ** 0.5))
dm = C.c_void_p()
_lib.check(lib.mag_devmesh_perforated(ctx.handle, nx, ny, 2.0, 64, 16, 3.0, C.byref(dm)), "perforated")
view = _lib.MagMesh()
_lib.check(lib.mag_devmesh_view(dm, C.byref(view)), "view")
N, E = int(view.n_nodes), int(view.n_elems)
m = meshgen.EXAMPLE_MATERIAL
mat = _lib.MagMaterial(m.youngs_modulus, m.poisson_ratio, m.part_thickness)
out = [torch.empty(N, dtype=torch.float64, device="cuda") for _ in range(4)] + [torch.empty(E, dtype=torch.float64, device="cuda")]
res = _lib.MagResult(*(t.data_ptr() for t in out), None, 1)
opt = _lib.default_options(precond=2)
sysh = C.c_void_p(); st = _lib.MagStats()
_lib.check(lib.mag_assemble(ctx.handle, C.byref(view), C.byref(mat), C.byref(opt), C.byref(sysh), C.byref(st)), "assemble")
for rep in range(2):
    s = _lib.MagStats()
    _lib.check(lib.mag_system_solve(sysh, C.byref(opt), C.byref(res), C.byref(s)), "solve")
    if rank == 0:
        print(f"C5 two-level world {world}: grid {nx}x{ny}, {E} triangles, {s.n_free} free DOF, n_coarse {s.n_coarse}, rep {rep}: "
              f"{s.iters} iterations, solve {s.ms_solve / 1e3:.3f} s (coarse setup {s.ms_coarse_setup / 1e3:.3f} s), "
              f"{1e3 * s.ms_solve / max(s.iters, 1):.1f} us/it, rel.res {s.final_residual / s.b_norm:.2e}, "
              f"ux range [{float(out[0].min()):.3f}, {float(out[0].max()):.3f}]", flush=True)
lib.mag_system_free(sysh)
lib.mag_devmesh_free(dm)
if world > 1:
    import torch.distributed as dist
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
else:
    ctx.close()
