#!/bin/bash
# Builds the reference's own binaries (dvbs2_tx_rx_bb, dvbs2_tx, dvbs2_ch, dvbs2_rx*) twice from a COPY of /root/reference:
#   baseline/_ref/bin/cpu/  unmodified reference (the chain CPU baseline from the real binary)
#   baseline/_ref/bin/gpu/  the same tree with baseline/apply_gpu_patch.py applied (--dec-hw GPU)
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
TARGETS="dvbs2_tx_rx_bb dvbs2_tx dvbs2_ch dvbs2_rx"

if [ "$WHAT" = cpu ] || [ "$WHAT" = all ]; then
    if [ -f "$W/src/.gpu_patched" ]; then # the CPU binaries come from the UNMODIFIED tree only: start over from a fresh copy
        rm -rf "$W/src" "$W/build_cpu" && cp -r "$REF" "$W/src" && chmod -R u+w "$W/src"
    fi
    configure "$W/src" "$W/build_cpu" -DDVBS2_LINK_CUDA=OFF
    make -C "$W/build_cpu" -j"$JOBS" $TARGETS > "$W/build_cpu.make.log" 2>&1
    for t in $TARGETS; do cp "$W/build_cpu/bin/$t" "$OUT/bin/cpu/"; strip "$OUT/bin/cpu/$t"; done
fi

if [ "$WHAT" = gpu ] || [ "$WHAT" = all ]; then
    # The AFF3CT / StreamPU libraries do not change, so the GPU binaries are built in the SAME build directory after the CPU
    # binaries have been copied out: patch the tree copy in place (once), reconfigure with the GPU option, rebuild the mains.
    [ -d "$W/build_cpu" ] || { echo "build the cpu binaries first"; exit 1; }
    if [ ! -f "$W/src/.gpu_patched" ]; then
        python "$HERE/apply_gpu_patch.py" "$W/src" "$ROOT" && touch "$W/src/.gpu_patched"
    fi
    configure "$W/src" "$W/build_cpu" -DDVBS2_LINK_CUDA=ON -DDVBS2_CUDA_ROOT="$ROOT"
    GPU_TARGETS="dvbs2_tx_rx_bb dvbs2_rx"
    make -C "$W/build_cpu" -j"$JOBS" $GPU_TARGETS > "$W/build_gpu.make.log" 2>&1
    for t in $GPU_TARGETS; do cp "$W/build_cpu/bin/$t" "$OUT/bin/gpu/"; strip "$OUT/bin/gpu/$t"; done
fi
rm -rf "$OUT/conf" && cp -r "$W/src/conf" "$OUT/conf"
mkdir -p "$OUT/run"   # binaries open ../conf/...: run them with cwd = baseline/_ref/run
echo "built $(ls "$OUT"/bin/*/ | wc -l) binaries under $OUT/bin"
