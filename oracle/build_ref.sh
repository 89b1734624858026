#!/bin/bash
# oracle/build_ref.sh X Y [strict|fast] — TEST INFRASTRUCTURE ONLY.
#
# Compiles the UNMODIFIED reference (cgmb/euler) from the sources where they lie under
# $REF into oracle/_ref/libeuler_ref_<X>x<Y>[_fast].so.  Nothing is copied into this repo:
# main.c is piped through sed — only the two enum lines that fix the grid size
# (main.c:23-24) are rewritten, inside the pipe — straight into gcc's stdin.  `main` is
# renamed with -D so the result is a dlopen-able library in which every function and
# global of main.c (all have external linkage) can be reached with ctypes.
#
#   strict: -O2 -ffp-contract=off, no fast-math — bit-identical to -O0;
#           this is the PARITY oracle.
#   fast  : the reference's own Release flags (CMakeLists.txt:11,18: -O3 -ffast-math and
#           -march=...) — this is the CPU TIMING baseline.  -march=native is replaced by
#           x86-64-v3 (AVX2+FMA) because the .so is built in the build container and run
#           on the GPU box's host CPU, which may be a different model.
#
# The reference's own build system (cmake) is not used.
set -euo pipefail
X=$1; Y=$2; FLAVOR=${3:-strict}
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)
OUT=$HERE/_ref
if [ ! -f "$REF/main.c" ]; then
  echo "build_ref: $REF/main.c not present; keeping prebuilt $OUT" >&2; exit 0
fi
mkdir -p "$OUT"
case $FLAVOR in
  strict) FLAGS="-std=gnu99 -O2 -DNDEBUG -ffp-contract=off"; SUF="";;
  fast)   FLAGS="-std=gnu99 -O3 -DNDEBUG -ffast-math -march=x86-64-v3"; SUF="_fast";;
  *) echo "flavor must be strict|fast" >&2; exit 2;;
esac
MCMODEL=""
# static data > 2 GB needs the medium code model
if [ $((X*Y)) -gt 16000000 ]; then MCMODEL="-mcmodel=medium"; fi
LIB=$OUT/libeuler_ref_${X}x${Y}${SUF}.so
sed -e "s/^  X = 100,/  X = ${X},/" -e "s/^  Y = 40\$/  Y = ${Y}/" "$REF/main.c" | \
  gcc $FLAGS $MCMODEL -fPIC -shared -w -Dmain=euler_tty_main -I"$REF" \
      -x c - "$REF"/misc/{terminal,file,rng,debug,time}.c -lm -o "$LIB"
# offsets of one exported global and of randf()'s function-local static RNG state, so the
# test harness can save/restore the RNG stream (it has no other handle on it).
nm "$LIB" | grep -E ' (g_u|rng_state[.0-9]*)$' > "${LIB%.so}.syms"
echo "built $LIB"
