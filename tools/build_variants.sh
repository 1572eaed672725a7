#!/bin/bash
# Builds A/B variants of the library into variants/ (git-ignored, travels to the GPU box; selected with AECM_B200_LIB):
#   v7.so          the round-1 kernel (one CTA per 28 rows, no row loop), from git
#   v8_free.so     the persistent kernel without the CTA-wide rendezvous between rows (-DAECM_ROW_RESYNC=0)
#   v8_resync2.so  rendezvous only before every 2nd row (-DAECM_ROW_RESYNC=2)
#   v8_syncframe/_syncblock.so  additional rendezvous before every frame / every potential block (-DAECM_SYNC_POINTS=1 / 2)
#   v8_cta14.so    two 14-warp CTAs per SM instead of one 28-warp CTA (-DAECM_WARPS_PER_CTA=14)
#   v8_nobin64.so  MEASUREMENT ONLY, wrong output: the one-live-lane pass for bin 64 compiled out (-DAECM_BIN_PASSES=2)
#   v8_fma1/2/3.so FFT butterflies as opaque multiply-add chains, alu-pipe work moved to the fma pipe: forward only, inverse only,
#                  both (-DAECM_FMA_BIAS=1 / 2 / 3)
# usage: tools/build_variants.sh [name ...]   (no names = all of them)
set -e
WANT=" $* "
want() { [ "$WANT" = "  " ] || [[ "$WANT" == *" $1 "* ]]; }
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/variants"
build() {  # name, source tree, extra nvcc flags
  want "$1" || return 0
  make -s -C "$2/webrtc_aecm_b200/csrc" clean >/dev/null 2>&1 || true
  make -s -C "$2/webrtc_aecm_b200/csrc" NVEXTRA="$3" "$2/webrtc_aecm_b200/libaecm_b200.so"
  cp "$2/webrtc_aecm_b200/libaecm_b200.so" "$ROOT/variants/$1.so"
}
TMP=$(mktemp -d)
git -C "$ROOT" archive 7f30022 webrtc_aecm_b200/csrc include | tar -x -C "$TMP"
build v7 "$TMP" ""
rm -rf "$TMP"; TMP=$(mktemp -d)
(cd "$ROOT" && tar -c webrtc_aecm_b200/csrc/*.c webrtc_aecm_b200/csrc/*.cu webrtc_aecm_b200/csrc/*.cuh webrtc_aecm_b200/csrc/*.h webrtc_aecm_b200/csrc/Makefile include tools/scalar_api_bench.c) | tar -x -C "$TMP"
build v8_free "$TMP" "-DAECM_ROW_RESYNC=0"
build v8_cta14 "$TMP" "-DAECM_WARPS_PER_CTA=14"
build v8_resync2 "$TMP" "-DAECM_ROW_RESYNC=2"
build v8_syncframe "$TMP" "-DAECM_SYNC_POINTS=1"
build v8_syncblock "$TMP" "-DAECM_SYNC_POINTS=2"
build v8_nobin64 "$TMP" "-DAECM_BIN_PASSES=2"
build v8_fma1 "$TMP" "-DAECM_FMA_BIAS=1"
build v8_fma2 "$TMP" "-DAECM_FMA_BIAS=2"
build v8_fma3 "$TMP" "-DAECM_FMA_BIAS=3"
rm -rf "$TMP"
ls -la "$ROOT/variants"
