"""SpMV time and algorithmic bandwidth of the device formats on a device-generated plate.

    python profiles/spmv_format_probe.py [nx ny]
"""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from magnetite_b200 import _lib, meshgen  # noqa: E402

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4000, 2000)
lib = _lib.load()
ctx = _lib.Context(0)
dm = C.c_void_p()
_lib.check(lib.mag_devmesh_plate(ctx.handle, nx, ny, 2.0, 3.0, C.byref(dm)), "plate")
view = _lib.MagMesh()
_lib.check(lib.mag_devmesh_view(dm, C.byref(view)), "view")
m = meshgen.EXAMPLE_MATERIAL
mat = _lib.MagMaterial(m.youngs_modulus, m.poisson_ratio, m.part_thickness)
for label, build_fmt, run_fmt in (("SELL-32, 32-bit columns", 3, 2), ("SELL-32, packed 16-bit offsets (default)", 0, 2),
                                  ("scalar CSR, thread per row", 0, 1)):
    sysh = C.c_void_p(); st = _lib.MagStats()
    _lib.check(lib.mag_assemble(ctx.handle, C.byref(view), C.byref(mat), C.byref(_lib.default_options(spmv_format=build_fmt)),
                                C.byref(sysh), C.byref(st)), "assemble")
    ms, nb = C.c_float(), C.c_uint64()
    for rep in range(2):
        _lib.check(lib.mag_system_spmv_bench(sysh, run_fmt, 100, C.byref(ms), C.byref(nb)), "bench")
    print(f"plate {nx}x{ny}  {label:36s}: {ms.value * 1e3:7.1f} us/SpMV, {nb.value / 1e9:6.3f} GB algorithmic, "
          f"{nb.value / (ms.value * 1e-3) / 1e9:7.1f} GB/s", flush=True)
    lib.mag_system_free(sysh)
lib.mag_devmesh_free(dm)
ctx.close()
