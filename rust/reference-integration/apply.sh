#!/bin/bash
# Puts libmagnetite_b200.so behind the reference's own entry points:
#     rust/reference-integration/apply.sh /path/to/Magnetite
#  1. copies solver_b200.rs into the checkout's src/ (it names crate::datatypes / crate::error, so it has to be
#     a module of the reference's binary crate);
#  2. adds the magnetite-b200-sys dependency (path = this repository's rust/magnetite-b200-sys) to Cargo.toml;
#  3. src/main.rs: `mod solver_b200;`, and the two call sites main.rs:64 (solver::run) and main.rs:69
#     (post_processor::csv_output) switch to solver_b200::run / solver_b200::csv_output.
# `mod solver;` stays: mesher::check_ccw keeps calling solver::compute_element_area (mesher.rs:9,523).
# Every edit is checked; nothing is left half-applied silently.  Then:
#     make -C <this repo>/magnetite_b200/csrc && cargo build --release
#     LD_LIBRARY_PATH=<this repo>/magnetite_b200 ./target/release/magnetite input.json geometry.svg
# UNVERIFIED beyond the text edits: no Rust toolchain exists in this repository's build image.
set -euo pipefail
REF=${1:?usage: apply.sh /path/to/Magnetite-checkout}
HERE=$(cd "$(dirname "$0")" && pwd)
SYS=$(cd "$HERE/../magnetite-b200-sys" && pwd)
MAIN=$REF/src/main.rs
TOML=$REF/Cargo.toml
for f in "$MAIN" "$TOML" "$REF/src/datatypes.rs" "$REF/src/error.rs"; do
    [ -f "$f" ] || { echo "apply.sh: $f not found — not a Magnetite checkout?" >&2; exit 1; }
done
if grep -q 'solver_b200' "$MAIN"; then echo "apply.sh: already applied" >&2; exit 1; fi
grep -q '^mod solver;' "$MAIN"                       || { echo "apply.sh: 'mod solver;' not found in main.rs" >&2; exit 1; }
grep -q 'solver::run(&mut nodes' "$MAIN"             || { echo "apply.sh: the solver::run call site not found" >&2; exit 1; }
grep -q 'post_processor::csv_output(' "$MAIN"        || { echo "apply.sh: the csv_output call site not found" >&2; exit 1; }
grep -q '^\[dependencies\]' "$TOML"                  || { echo "apply.sh: [dependencies] not found in Cargo.toml" >&2; exit 1; }

cp "$HERE/solver_b200.rs" "$REF/src/solver_b200.rs"
sed -i "s|^\[dependencies\]|[dependencies]\nmagnetite-b200-sys = { path = \"$SYS\" }|" "$TOML"
sed -i -e 's|^mod solver;|mod solver;\nmod solver_b200;|' \
       -e 's|solver::run(&mut nodes|solver_b200::run(\&mut nodes|' \
       -e 's|post_processor::csv_output(|solver_b200::csv_output(|' "$MAIN"

grep -q '^mod solver_b200;' "$MAIN" && grep -q 'solver_b200::run(&mut nodes' "$MAIN" \
    && grep -q 'solver_b200::csv_output(' "$MAIN" && grep -q '^magnetite-b200-sys' "$TOML" \
    || { echo "apply.sh: an edit did not take" >&2; exit 1; }
echo "apply.sh: $REF now calls libmagnetite_b200.so (solver_b200::run, solver_b200::csv_output)"
