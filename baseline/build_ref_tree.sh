#!/bin/bash
# Builds the reference's own binaries (dvbs2_tx_rx_bb, dvbs2_tx, dvbs2_ch, dvbs2_rx*) twice from a COPY of /root/reference:
#   baseline/_ref/bin/cpu/  unmodified reference (the chain CPU baseline from the real binary)
#   baseline/_ref/bin/gpu/  the same tree with baseline/dvbs2_gpu.patch applied (--dec-hw GPU)
# The copy and the object files live in baseline/_build/ (git- and gpurun-ignored); only binaries + conf/ go to
# baseline/_ref/ (git-ignored, travels to the GPU box).  Recipe = SURVEY.md App. C.  No reference source enters the repo.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
W="$HERE/_build"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"
WHAT="${1:-all}"   # cpu | gpu | all
[ -d "$REF" ] || { echo "no reference tree at $REF"; exit 0; }
mkdir -p "$W" "$OUT/bin/cpu" "$OUT/bin/gpu"

configure() { # $1 = source dir, $2 = build dir, rest = extra cmake args
    local S="$1" B="$2"; shift 2
    cmake -S "$S" -B "$B" -G"Unix Makefiles" -DCMAKE_BUILD_TYPE=Release \
        -DCMAKE_CXX_FLAGS="-funroll-loops -march=x86-64-v4" -DDVBS2_LINK_UHD=OFF -DSPU_LINK_HWLOC=OFF \
        -DSPU_OVERRIDE_VERSION=v1.6.1 -DAFF3CT_OVERRIDE_VERSION=v3.0.2 -DSPU_STACKTRACE=OFF "$@" > "$B.configure.log" 2>&1
}

if [ ! -d "$W/src" ]; then
    cp -r "$REF" "$W/src" && chmod -R u+w "$W/src"
fi
TARGETS="dvbs2_tx_rx_bb dvbs2_tx dvbs2_ch dvbs2_rx dvbs2_rx8"

if [ "$WHAT" = cpu ] || [ "$WHAT" = all ]; then
    configure "$W/src" "$W/build_cpu"
    make -C "$W/build_cpu" -j"$JOBS" $TARGETS > "$W/build_cpu.make.log" 2>&1
    for t in $TARGETS; do cp "$W/build_cpu/bin/$t" "$OUT/bin/cpu/"; done
fi

if [ "$WHAT" = gpu ] || [ "$WHAT" = all ]; then
    rm -rf "$W/src_gpu" && cp -r "$W/src" "$W/src_gpu"
    (cd "$W/src_gpu" && patch -p1 < "$HERE/dvbs2_gpu.patch")
    # the libraries do not change: reuse the CPU build's objects
    [ -d "$W/build_gpu" ] || cp -r "$W/build_cpu" "$W/build_gpu"
    configure "$W/src_gpu" "$W/build_gpu" -DDVBS2_LINK_CUDA=ON -DDVBS2_CUDA_ROOT="$ROOT"
    make -C "$W/build_gpu" -j"$JOBS" $TARGETS > "$W/build_gpu.make.log" 2>&1
    for t in $TARGETS; do cp "$W/build_gpu/bin/$t" "$OUT/bin/gpu/"; done
fi
rm -rf "$OUT/conf" && cp -r "$W/src/conf" "$OUT/conf"
mkdir -p "$OUT/run"   # binaries open ../conf/...: run them with cwd = baseline/_ref/run
echo "built $(ls "$OUT"/bin/*/ | wc -l) binaries under $OUT/bin"
