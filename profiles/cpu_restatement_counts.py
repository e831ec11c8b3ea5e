"""CPU: iteration count and true residual of the numpy / scipy restatement of the default solver (oracle/two_level.py)
on a plate — run here, without a GPU, to compare with the B200's bench line of the same plate.

    python profiles/cpu_restatement_counts.py 4000 2000               # the headline workload: 11 GB, about 8 minutes on 8 cores
    python profiles/cpu_restatement_counts.py 4000 2000 perforated    # BASELINE config 5 at its one-GPU size (holes: pitch 64, radius 16)
"""
import json
import resource
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from magnetite_b200 import meshgen          # noqa: E402
from oracle import two_level as T           # noqa: E402

nx, ny = int(sys.argv[1]), int(sys.argv[2])
perforated = len(sys.argv) > 3 and sys.argv[3] == "perforated"
t = time.time()
mesh = meshgen.perforated_plate(nx, ny, pitch=64, radius=16) if perforated else meshgen.plate(nx, ny)
S = T.reduced_system(mesh, meshgen.EXAMPLE_MATERIAL)
t_asm = time.time() - t
t = time.time()
x, it = T.pcg(S, 2)
t_cg = time.time() - t
res = float(np.linalg.norm(S.rhs - S.A @ x) / np.linalg.norm(S.rhs))
print(json.dumps({"plate": f"{nx}x{ny}" + (" perforated (pitch 64, radius 16)" if perforated else ""), "triangles": int(mesh.n_elems), "n_free": int(S.A.shape[0]), "nnz": int(S.A.nnz),
                  "grid": list(T.coarse_grid(S.A.shape[0], S.box)[:2]), "two_level_iters": it, "true_rel_residual": res,
                  "seconds_assemble": round(t_asm, 1), "seconds_pcg": round(t_cg, 1),
                  "peak_rss_gb": round(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, 1)}))
