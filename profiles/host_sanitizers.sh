#!/bin/bash
# Host code of the whole repo under AddressSanitizer + UndefinedBehaviorSanitizer (gcc 13), on CPU, no GPU needed:
#   * libmagnetite_b200.so: the host side of api.cu (the C ABI, DeviceHeap, partition / halo planning, option and
#     error handling), csv.cpp, reorder.cpp — device code is untouched by the flags;
#   * the oracle's C restatement; the C++ host layer binaries (host/magnetite_b200, host/plate_demo);
# then `pytest -m "not gpu"` against those builds, UBSan halting on the first report.  The normal builds are put
# back afterwards.  Result of the last run: profiles/r2_host_sanitizers.txt.
# (compute-sanitizer for the DEVICE code is closed on the GPU pool — DESIGN.md §4d.)
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/magnetite_b200/csrc
TMP=$(mktemp -d)
SAN="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"

cp "$ROOT/magnetite_b200/libmagnetite_b200.so" "$TMP/lib.keep"
cp "$ROOT/oracle/_build/libmagnetite_oracle.so" "$TMP/orc.keep"
cp "$ROOT/host/magnetite_b200" "$TMP/cli.keep"
cp "$ROOT/host/plate_demo" "$TMP/demo.keep"
restore() {
    cp "$TMP/lib.keep" "$ROOT/magnetite_b200/libmagnetite_b200.so"
    cp "$TMP/orc.keep" "$ROOT/oracle/_build/libmagnetite_oracle.so"
    cp "$TMP/cli.keep" "$ROOT/host/magnetite_b200"
    cp "$TMP/demo.keep" "$ROOT/host/plate_demo"
    rm -rf "$TMP"
}
trap restore EXIT

$NVCC -O1 -g -std=c++17 $ARCH --expt-relaxed-constexpr -I/usr/include -I"$SRC" \
    -Xcompiler -fPIC,-ffp-contract=off,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer \
    -c -o "$TMP/api.o" "$SRC/api.cu"
g++ $SAN -std=c++17 -fPIC -c -o "$TMP/csv.o" "$SRC/csv.cpp"
g++ $SAN -std=c++17 -fPIC -c -o "$TMP/reorder.o" "$SRC/reorder.cpp"
$NVCC -shared -o "$ROOT/magnetite_b200/libmagnetite_b200.so" "$TMP/api.o" "$TMP/csv.o" "$TMP/reorder.o" $ARCH \
    -L/usr/lib/x86_64-linux-gnu -l:libnccl.so.2 -lcudart -Xcompiler -fsanitize=address,-fsanitize=undefined
gcc $SAN -ffp-contract=off -fPIC -std=c11 -D_POSIX_C_SOURCE=200809L -shared \
    -o "$ROOT/oracle/_build/libmagnetite_oracle.so" "$ROOT/oracle/magnetite_oracle.c" -lm
LINK="-L$ROOT/magnetite_b200 -lmagnetite_b200 -L/usr/local/cuda/lib64 -Wl,-rpath,$ROOT/magnetite_b200 -Wl,-rpath,/usr/local/cuda/lib64 -Wl,--allow-shlib-undefined"
( cd "$ROOT/host"
  g++ $SAN -std=c++17 -ffp-contract=off -o magnetite_b200 magnetite_cli.cpp magnetite_io.cpp magnetite_host.cpp $LINK
  g++ $SAN -std=c++17 -ffp-contract=off -o plate_demo plate_demo.cpp magnetite_host.cpp $LINK )

cd "$ROOT"
LD_PRELOAD=/usr/lib/x86_64-linux-gnu/libasan.so.8:/usr/lib/x86_64-linux-gnu/libubsan.so.1 \
ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 \
    python -m pytest tests -q -m "not gpu" -p no:cacheprovider
