#!/bin/bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference (bbuhrow/avx-ecm) from the sources where they
# lie under $REF (default /root/reference) into oracle/_ref/.  Nothing is copied
# into the repo; oracle/_ref/ is git-ignored (but travels to the GPU box).
#
#   avx-ecm-ref     DIGITBITS=52, 8 curves per vector  (SKYLAKEX flag set, Makefile:88-91)
#   avx-ecm-ref32   DIGITBITS=32, 16 curves per vector (used for the R-independence test)
#
# Recipe notes: no gmp.h in the image -> declarations-only
# shim + libgmp.so.10 by path; -fcommon for the tentative definitions in the
# reference headers; -ffp-contract=off so prac()'s double arithmetic is not fused.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
GMPLIB="${GMPLIB:-/usr/lib/x86_64-linux-gnu/libgmp.so.10}"
if [ ! -d "$REF" ]; then echo "build_ref: $REF absent, keeping prebuilt $OUT" ; exit 0; fi
if ! grep -q avx512f /proc/cpuinfo || ! grep -q avx512dq /proc/cpuinfo; then
  echo "build_ref: host lacks AVX-512F/DQ, reference cannot be built (no portable build exists)"; exit 0; fi
mkdir -p "$OUT"
SRCS="eratosthenes/presieve.c eratosthenes/count.c eratosthenes/offsets.c eratosthenes/primes.c
 eratosthenes/roots.c eratosthenes/linesieve.c eratosthenes/soe.c eratosthenes/tiny.c
 eratosthenes/worker.c eratosthenes/soe_util.c eratosthenes/wrapper.c threadpool.c main.c ecm.c
 util.c vecarith.c vecarith52.c vec_common.c calc.c queue.c"
FILES=""
for s in $SRCS; do FILES="$FILES $REF/$s"; done
CFLAGS="-fcommon -O3 -g0 -mavx -march=skylake-avx512 -DSKYLAKEX -ffp-contract=off -w -I$HERE/shim -I$REF -I$REF/eratosthenes"
build() { # name extra-flags
  if [ -x "$OUT/$1" ] && [ "$OUT/$1" -nt "$REF/ecm.c" ] && [ "$OUT/$1" -nt "$0" ]; then return; fi
  gcc $CFLAGS $2 $FILES "$HERE/shim/ref_stub.c" -o "$OUT/$1" "$GMPLIB" -lm -lpthread
  echo "built $OUT/$1"
}
# fieldop-ref: the reference's special-form field operators behind a one-op-per-line driver
# (shim/fieldop_harness.c); the reference's main() is renamed on the command line, not in its source
build_harness() {
  if [ -x "$OUT/fieldop-ref" ] && [ "$OUT/fieldop-ref" -nt "$REF/vecarith52.c" ] && [ "$OUT/fieldop-ref" -nt "$HERE/shim/fieldop_harness.c" ]; then return; fi
  T="$(mktemp -d)"
  ( cd "$T" && gcc $CFLAGS -Dmain=ref_main -c $FILES "$HERE/shim/ref_stub.c" && gcc $CFLAGS -c "$HERE/shim/fieldop_harness.c" -o harness_main.o \
    && gcc *.o -o "$OUT/fieldop-ref" "$GMPLIB" -lm -lpthread ) && echo "built $OUT/fieldop-ref"
  rm -rf "$T"
}
build avx-ecm-ref "" &
build avx-ecm-ref32 "-DDIGITBITS=32" &
build_harness &
wait
if [ ! -f "$OUT/gcd_tap.so" ] || [ "$HERE/shim/gcd_tap.c" -nt "$OUT/gcd_tap.so" ]; then
  gcc -O2 -fPIC -shared -I"$HERE/shim" "$HERE/shim/gcd_tap.c" -o "$OUT/gcd_tap.so" -ldl "$GMPLIB"
fi
