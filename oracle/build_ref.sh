#!/usr/bin/env bash
# Build oracle/_ref/libref_hashgrid.so from the reference's own kernel text.
#
# Compiles /root/reference/src/encoder/hashencoder/src/hashencoder.cu lines 30-298
# (the templates: div_round_up, fast_hash, get_grid_index, kernel_grid,
# kernel_grid_backward, kernel_input_backward -- everything above line 30 is ATen/CUDA
# includes and an at::Half atomic, everything below 298 is the launch/ATen glue) as host
# C++ behind oracle/ref_shim.cpp.  The extracted text goes to a temp dir OUTSIDE the repo
# and is deleted afterwards; only the .so lands in oracle/_ref/ (git-ignored).
#
# -mfma -ffp-contract=fast makes g++ contract  x*scale+0.5f  and  res += w*grid  into FMAs
# exactly where nvcc (-fmad=true, the default the reference is JIT-built with) emits FFMA
# (SURVEY.md section 0 / 8c SASS probe).  tests/test_oracle_pin.py asserts the contraction
# really happened by comparing against nafb_oracle.c's explicit fmaf() bit for bit.
set -euo pipefail
REF_ROOT="${REF_ROOT:-/root/reference}"
SRC="$REF_ROOT/src/encoder/hashencoder/src/hashencoder.cu"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$SRC" ]; then
    echo "build_ref.sh: $SRC not found (no reference mount on this machine) -- skipping" >&2
    exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/nafb_ref.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
sed -n '30,298p' "$SRC" > "$TMP/ref_kernels.inc"
g++ -O2 -std=c++17 -mfma -ffp-contract=fast -fopenmp -fPIC -shared -w \
    -DREF_KERNEL_TEXT="\"$TMP/ref_kernels.inc\"" \
    -o "$OUT/libref_hashgrid.so" "$HERE/ref_shim.cpp" -lm
echo "built $OUT/libref_hashgrid.so from $SRC:30-298"
