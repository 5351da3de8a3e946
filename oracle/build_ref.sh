#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference path into oracle/_ref/ (git-ignored, travels with gpurun).
#
# Compiles, with plain g++ and no cmake, the reference sources where they lie under /root/reference:
#   * every StreamPU .cpp (runtime the AFF3CT modules derive from),
#   * the AFF3CT .cpp files the path needs (list in REF_AFF3CT below; discovered by link errors),
#   * dvbs2's own BCH wrapper / encoder,
#   * oracle/ref_harness.cpp (ours) which exposes them through a C ABI.
# Output: oracle/_ref/libdvbs2_ref.so.  Nothing is copied out of /root/reference.
# Flags follow the reference's release build (CMakeLists.txt:7,93): C++11 for library
# sources (so --dec-simd INTER resolves to ONMS_inter exactly as in the dvbs2 binaries), -O3 -funroll-loops,
# AFF3CT_MULTI_PREC.  -march is x86-64-v4 when the build host has AVX-512 (16 float lanes, the north-star
# "-F 16" baseline) and falls back to x86-64-v3.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OBJ=$OUT/obj
JOBS=${JOBS:-$(nproc)}
[ -d "$REF" ] || { echo "no reference tree at $REF — keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OBJ"

if grep -q avx512bw /proc/cpuinfo; then MARCH=${MARCH:-x86-64-v4}; else MARCH=${MARCH:-x86-64-v3}; fi
A=$REF/lib/aff3ct
S=$REF/lib/streampu
INC="-I$A/include -I$A/src -I$S/include -I$S/src -I$A/lib/MIPP/src -I$A/lib/cli/src -I$A/lib/date/include/date \
 -I$S/lib/rang/include -I$S/lib/json/include -I$REF/src/common"
DEF="-DAFF3CT_MULTI_PREC -DAFF3CT_EXT_STRINGS -DAFF3CT_POLAR_BIT_PACKING -DSPU_COLORS -DNDEBUG"
CXXFLAGS="-O3 -funroll-loops -march=$MARCH -fPIC -w $DEF $INC"

REF_AFF3CT="
src/Tools/Code/LDPC/Standard/DVBS2/DVBS2_constants.cpp
src/Tools/Algo/Matrix/Sparse_matrix/Sparse_matrix.cpp
src/Tools/Algo/Matrix/Matrix.cpp
src/Tools/Math/Galois.cpp
src/Tools/Code/BCH/BCH_polynomial_generator.cpp
src/Tools/general_utils.cpp
src/Tools/Noise/Noise.cpp
src/Tools/Noise/Sigma.cpp
src/Tools/Noise/Event_probability.cpp
src/Tools/Noise/Received_optical_power.cpp
src/Tools/Interleaver/NO/Interleaver_core_NO.cpp
src/Tools/Interleaver/Column_row/Interleaver_core_column_row.cpp
src/Module/Decoder/Decoder.cpp
src/Module/Decoder/LDPC/BP/Decoder_LDPC_BP.cpp
src/Module/Decoder/LDPC/BP/Horizontal_layered/ONMS/Decoder_LDPC_BP_horizontal_layered_ONMS_inter.cpp
src/Module/Decoder/BCH/Decoder_BCH.cpp
src/Module/Decoder/BCH/Standard/Decoder_BCH_std.cpp
src/Module/Encoder/BCH/Encoder_BCH.cpp
src/Module/Encoder/LDPC/Encoder_LDPC.cpp
src/Module/Encoder/LDPC/DVBS2/Encoder_LDPC_DVBS2.cpp
src/Module/Quantizer/Pow2/Quantizer_pow2.cpp
src/Tools/Perf/common/hard_decide.cpp
"
[ -n "${REF_EXTRA:-}" ] && REF_AFF3CT="$REF_AFF3CT $REF_EXTRA"

list=$OBJ/units.txt; : > "$list"
for f in $(cd "$S" && find src -name '*.cpp' | sort); do echo "$S/$f spu_$(echo "$f" | tr '/' '_' | sed 's/\.cpp$//').o 11" >> "$list"; done
for f in $REF_AFF3CT; do echo "$A/$f aff_$(echo "$f" | tr '/' '_' | sed 's/\.cpp$//').o 11" >> "$list"; done
for f in Module/Decoder_BCH_DVBS2/Decoder_BCH_DVBS2.cpp Module/Encoder_BCH_DVBS2/Encoder_BCH_DVBS2.cpp; do
  echo "$REF/src/common/$f dvb_$(echo "$f" | tr '/' '_' | sed 's/\.cpp$//').o 11" >> "$list"; done
echo "$HERE/ref_harness.cpp harness.o 17" >> "$list"

compile_one() { # src obj std
  local src=$1 obj=$2 std=$3
  if [ ! -f "$OBJ/$obj" ] || [ "$src" -nt "$OBJ/$obj" ]; then
    g++ -std=c++$std $CXXFLAGS -c "$src" -o "$OBJ/$obj" || { echo "FAILED: $src" >&2; return 1; }
  fi
}
export -f compile_one; export OBJ CXXFLAGS
xargs -P "$JOBS" -L 1 bash -c 'compile_one "$0" "$1" "$2"' < "$list"

g++ -shared -o "$OUT/libdvbs2_ref.so" $(awk -v o="$OBJ" '{print o"/"$2}' "$list") -lpthread -Wl,--no-undefined
echo "march=$MARCH" > "$OUT/BUILD_INFO"
echo "built $OUT/libdvbs2_ref.so"
