"""The two assembly modes of mag_options.assembly on the device-resident plate — 0 gather (default), 1 sorted COO
keys + segmented reduction; both followed by the same elimination kernels:
    python profiles/assembly_probe.py [nx ny reps [modes]]   (default 4000 2000 3 01 = 16 M DOF, both modes)

Prints the library's phase timers (CUDA events on its stream) for every repetition and checks at full size
that all paths leave the same K_ff: same counts, and the order-preserving CSR SpMV of the systems on one
random vector gives bit-identical results."""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from magnetite_b200 import _lib, meshgen, solver  # noqa: E402

PHASES = ("ms_upload", "ms_elem", "ms_sort", "ms_reduce", "ms_bc", "ms_format", "ms_total")


def main():
    nx, ny, reps = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4000, 2000, 3)
    modes = [int(c) for c in (sys.argv[4] if len(sys.argv) > 4 else "01")]
    names = {0: "gather", 1: "sorted keys"}
    lib = _lib.load()
    ctx = _lib.Context(0)
    mat = solver._material(meshgen.EXAMPLE_MATERIAL)

    def plate(px, py):
        dm, view = C.c_void_p(), _lib.MagMesh()
        _lib.check(lib.mag_devmesh_plate(ctx.handle, px, py, 2.0, 3.0, C.byref(dm)), "mag_devmesh_plate")
        _lib.check(lib.mag_devmesh_view(dm, C.byref(view)), "mag_devmesh_view")
        return dm, view

    def assemble(view, assembly):
        opt = _lib.default_options(assembly=assembly)
        sysh, st = C.c_void_p(), _lib.MagStats()
        _lib.check(lib.mag_assemble(ctx.handle, C.byref(view), C.byref(mat), C.byref(opt), C.byref(sysh), C.byref(st)),
                   f"mag_assemble(assembly={assembly})")
        return sysh, st.as_dict()

    wdm, wview = plate(256, 128)                       # warm-up: module load, heap slabs
    for a in modes:
        s, _ = assemble(wview, a)
        lib.mag_system_free(s)
    lib.mag_devmesh_free(wdm)

    dm, view = plate(nx, ny)
    n_elems = int(view.n_elems)
    keep = {}
    for a in modes:
        for rep in range(reps):
            s, st = assemble(view, a)
            asm_ms = sum(st[k] for k in ("ms_elem", "ms_sort", "ms_reduce", "ms_bc"))
            row = {"assembly": names[a], "rep": rep, "elements": n_elems,
                   "assembly_ms": round(asm_ms, 3), "melem_per_s": round(n_elems / asm_ms / 1e3, 1),
                   "n_free": int(st["n_free"]), "nnz": int(st["nnz"]), "nnz_structural": int(st["nnz_structural"]),
                   "launches": int(st["kernel_launches"])}
            row.update({k: round(float(st[k]), 3) for k in PHASES})
            print(json.dumps(row), flush=True)
            if rep == reps - 1:
                keep[a] = (s, st)
            else:
                lib.mag_system_free(s)
    n_free = int(keep[modes[0]][1]["n_free"])
    x = np.random.default_rng(1).normal(size=n_free)
    ys = {}
    for a in modes:
        ys[a] = np.empty(n_free)
        _lib.check(lib.mag_system_spmv(keep[a][0], 1, _lib.ptr(x), _lib.ptr(ys[a])), f"spmv({names[a]})")
    st0 = keep[modes[0]][1]
    same_counts = all(int(st0[k]) == int(keep[a][1][k]) for a in modes for k in ("n_free", "nnz", "nnz_structural", "sell_entries"))
    print(json.dumps({"modes": [names[a] for a in modes], "same_counts": bool(same_counts),
                      "csr_spmv_bit_identical": bool(all(np.array_equal(ys[modes[0]], ys[a]) for a in modes)),
                      "spmv_norm": float(np.linalg.norm(ys[modes[0]]))}), flush=True)
    for a in modes:
        lib.mag_system_free(keep[a][0])
    lib.mag_devmesh_free(dm)
    ctx.close()


if __name__ == "__main__":
    main()
