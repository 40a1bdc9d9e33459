#!/bin/bash
# oracle/build_ref_surfdisp.sh -- compiles the reference's surfmodes/surfdisp96.f with a real Fortran compiler into
# oracle/_ref/surfdisp96_gfortran, a filter that reads columns from stdin (format of ref_harness/surfdisp_driver.f90)
# and writes surfdisp96's / surfdisp_mmodes' outputs.  Flags follow the reference: no optimisation flags for the .f file
# (src/makefile:154-155 uses the empty $(F77FLAGS)), plus -ffixed-line-length-0 without which lines beyond column 72
# (e.g. surfdisp96.f:286) do not parse; -O3 for the free-form driver like $(FFLAGS) (src/makefile:57-60).
# The build image has no Fortran compiler: tests/test_oracle_vs_fortran.py is REPORTED SKIPPED there and runs wherever
# gfortran is on PATH.  Nothing is copied from the reference; outputs go to the git-ignored oracle/_ref/ only.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MCT_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
FC="${FC:-gfortran}"
command -v "$FC" >/dev/null 2>&1 || { echo "build_ref_surfdisp: no $FC on PATH, skipping"; exit 0; }
[ -f "$REF/surfmodes/surfdisp96.f" ] || { echo "build_ref_surfdisp: $REF/surfmodes/surfdisp96.f not present, skipping"; exit 0; }
mkdir -p "$OUT"
TMP="$(mktemp -d)"
"$FC" -ffixed-line-length-0 -c "$REF/surfmodes/surfdisp96.f" -o "$TMP/surfdisp96.o"
"$FC" -O3 -o "$OUT/surfdisp96_gfortran" "$HERE/ref_harness/surfdisp_driver.f90" "$TMP/surfdisp96.o"
rm -rf "$TMP"
echo "build_ref_surfdisp: built $OUT/surfdisp96_gfortran"
