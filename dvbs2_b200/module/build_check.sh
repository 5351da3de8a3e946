#!/usr/bin/env bash
# Compiles module_check (the C++ drop-in test of the GPU modules) against the reference's headers where they lie under
# /root/reference, linking the compiled reference (oracle/_ref/libdvbs2_ref.so) and libdvbs2_cuda.so.
# Output: dvbs2_b200/module/_build/module_check (git-ignored; travels to the GPU box with gpurun).
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
[ -d "$REF" ] || { echo "no reference tree at $REF — keeping prebuilt module_check" >&2; exit 0; }
[ -f "$ROOT/oracle/_ref/libdvbs2_ref.so" ] || bash "$ROOT/oracle/build_ref.sh"
A=$REF/lib/aff3ct
S=$REF/lib/streampu
INC="-I$A/include -I$A/src -I$S/include -I$S/src -I$A/lib/MIPP/src -I$A/lib/cli/src -I$A/lib/date/include/date \
 -I$S/lib/rang/include -I$S/lib/json/include -I$REF/src/common -I$ROOT/include -I$HERE"
DEF="-DAFF3CT_MULTI_PREC -DAFF3CT_EXT_STRINGS -DAFF3CT_POLAR_BIT_PACKING -DSPU_COLORS -DNDEBUG"
MARCH=$(cut -d= -f2 "$ROOT/oracle/_ref/BUILD_INFO" 2>/dev/null || echo x86-64-v3)
mkdir -p "$HERE/_build"
g++ -std=c++17 -O2 -march=$MARCH -w $DEF $INC "$HERE/module_check.cpp" -o "$HERE/_build/module_check" \
  -L"$ROOT/oracle/_ref" -ldvbs2_ref -L"$ROOT/dvbs2_b200" -ldvbs2_cuda -lpthread \
  -Wl,-rpath,'$ORIGIN/../../../oracle/_ref' -Wl,-rpath,'$ORIGIN/../..'
echo "built $HERE/_build/module_check"
