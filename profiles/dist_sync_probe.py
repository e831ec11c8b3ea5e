"""torchrun worker: per-iteration time of the distributed PCG for several plate sizes and both
allreduce modes (0 = peer-memory mailbox, 1 = NCCL).  Small plates expose the synchronisation cost."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from magnetite_b200 import _lib, dist as mdist, meshgen, solver  # noqa: E402

import os
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    rank, world, local = mdist.init_process_group("nccl")
    ctx = _lib.Context(local)
    mdist.init_comm(ctx)
else:
    rank, world, local = 0, 1, 0
    ctx = _lib.Context(0)
meta = meshgen.EXAMPLE_MATERIAL
PRECOND = int(os.environ.get("MAG_PROBE_PRECOND", "1"))
for nx, ny in ((200, 100), (1000, 500), (2000, 1000)):
    mesh = meshgen.plate(nx, ny)
    for ar in ((0, 1) if world > 1 else (0,)):
        with solver.System(mesh, meta, ctx) as S:
            for rep in range(2):
                sol = S.solve(_lib.default_options(allreduce=ar, max_iter=2000, precond=PRECOND), allow_not_converged=True)
            st = sol.stats
            if rank == 0:
                pr = ", ".join(f"{v / 1e3:.1f}" for v in st["prof"][:7])
                print(f"precond {PRECOND} tune {os.environ.get('MAG_TUNE', '0')} world {world} plate {nx}x{ny} allreduce {ar}: {st['iters']} it, "
                      f"{1e3 * st['ms_solve'] / max(st['iters'], 1):.1f} us/it | timeline us "
                      f"[gapA, durA, gapB, waitB, durB, gapC, waitC] = [{pr}]", flush=True)
if world > 1:
    import torch.distributed as dist
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
else:
    ctx.close()
