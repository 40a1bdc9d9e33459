#!/bin/bash
# oracle/build_ref.sh -- builds the reference-derived checkers under the git-ignored oracle/_ref/:
#   libsurfdisp96_f2c.so  the reference's surfdisp96.f through oracle/f77toc.py (mechanical F77 -> C) + gcc
#   libfm2d_ttime_f2c.so  the reference's fm2d/fm2d_ttime.f90 through oracle/f90toc.py (mechanical F90 -> C) + gcc
#   surfdisp96_gfortran   the same file compiled by gfortran, where one exists (oracle/build_ref_surfdisp.sh)
#   kdtree2_ref           from the reference's OWN pre-built
# object utils/libutils.a:kdtree2.o (linked where it lies; nothing is copied into the repo
# except the resulting binary under the git-ignored oracle/_ref/).  Needs /root/reference and a
# libgfortran.so.5 (the one bundled with scipy's wheel).  Exits 0 with a message if either is
# missing: the binary is optional test infrastructure.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MCT_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
# ---- stage 2: the reference's surfmodes/surfdisp96.f, translated mechanically to C (oracle/f77toc.py) and compiled
# with the strict-IEEE flags of the restatement.  No Fortran compiler is needed; the source is read where it lies.
if [ -f "$REF/surfmodes/surfdisp96.f" ]; then
  mkdir -p "$OUT"
  python "$HERE/f77toc.py" "$REF/surfmodes/surfdisp96.f" "$OUT/surfdisp96_f2c.c"
  gcc -O2 -fPIC -std=gnu11 -ffp-contract=off -fno-fast-math -shared -o "$OUT/libsurfdisp96_f2c.so" "$OUT/surfdisp96_f2c.c" -lm
  echo "build_ref: built $OUT/libsurfdisp96_f2c.so from $REF/surfmodes/surfdisp96.f"
else
  echo "build_ref: $REF/surfmodes/surfdisp96.f not present, skipping the translated surfdisp96"
fi
# ---- the fast-marching core fm2d/fm2d_ttime.f90 (+ the module variables of fm2d_globalp.f90), translated mechanically to C
# (oracle/f90toc.py) and compiled behind a small driver that stands in for modrays' allocations
if [ -f "$REF/fm2d/fm2d_ttime.f90" ] && [ -f "$REF/fm2d/fm2d_globalp.f90" ]; then
  mkdir -p "$OUT"
  python "$HERE/f90toc.py" "$REF/fm2d/fm2d_globalp.f90" "$REF/fm2d/fm2d_ttime.f90" \
      "$REF/fm2d/fm2dray_cartesian.f90:gridder,bsplrefine,srtimes,rpaths,@modrays_source" "$OUT/fm2d_ttime_f2c.c"
  gcc -O2 -fPIC -std=gnu11 -ffp-contract=off -fno-fast-math -shared -DFM2D_F2C_SOURCE="\"$OUT/fm2d_ttime_f2c.c\"" \
      -o "$OUT/libfm2d_ttime_f2c.so" "$HERE/ref_harness/fm2d_f90_harness.c" -lm
  echo "build_ref: built $OUT/libfm2d_ttime_f2c.so from $REF/fm2d/fm2d_ttime.f90 + gridder, bsplrefine, srtimes, rpaths and the body of the source loop of fm2dray_cartesian.f90"
else
  echo "build_ref: $REF/fm2d/fm2d_ttime.f90 not present, skipping the translated fast-marching core"
fi
# ---- the Love-wave generalized R/T secular function, surfmodes/Love.f90 (+ csq, parameters and T_GRT of GRT.f90), translated
# mechanically (oracle/f90toc_love.py); -fcx-fortran-rules: complex products and quotients as gfortran's middle end expands them
if [ -f "$REF/surfmodes/Love.f90" ] && [ -f "$REF/surfmodes/GRT.f90" ]; then
  mkdir -p "$OUT"
  python "$HERE/f90toc_love.py" "$REF/surfmodes/GRT.f90" "$REF/surfmodes/Love.f90" "$REF/surfmodes/util.f90:bisecim,sort" "$REF/surfmodes/C_interval_L.f90" \
      "$REF/surfmodes/Rayleigh.f90:startl:nohdr" "$REF/surfmodes/surfmodes.f90:setup_grt,calgroup" host=kt:real,c:real,grt:t_grt \
      "$REF/surfmodes/SearchLove.f90:check,fundamode" "$OUT/love_f2c.c"
  gcc -O2 -fPIC -std=gnu11 -fcx-fortran-rules -ffp-contract=off -fno-fast-math -shared -DLOVE_F2C_SOURCE="\"$OUT/love_f2c.c\"" \
      -o "$OUT/liblove_f2c.so" "$HERE/ref_harness/love_f90_harness.c" -lm
  echo "build_ref: built $OUT/liblove_f2c.so from $REF/surfmodes/Love.f90 + bisecim, sort of util.f90 + the C_Interval file"
  python "$HERE/f90toc_love.py" "$REF/surfmodes/GRT.f90" "$REF/surfmodes/Rayleigh.f90:inv2,init_rayleigh,delete_rayleigh,startl,secfunsurf,einve,propup,secfunst,stoneley,einve_f,propdn_f" "$REF/surfmodes/util.f90:bisecim,sort,det3" "$REF/surfmodes/C_interval.f90" "$REF/surfmodes/surfmodes.f90:setup_grt,calgroup" \
      "host=iq:integer,j:integer,ip:integer,ilay:integer,k1:real,k2:real,f1:real,f2:real,kt:real,c:real,imf:real,r:real,dr:real,v1:real*@cr0_finder,v2:real*@cr0_finder,grt:t_grt@st_finder" \
      "$REF/surfmodes/SearchRayleigh.f90:fundamode,stmode,cr0_finder,rayhomo,st_finder,getst" "$OUT/rayleigh_f2c.c"
  gcc -O2 -fPIC -std=gnu11 -fcx-fortran-rules -ffp-contract=off -fno-fast-math -shared -DRAYLEIGH_F2C_SOURCE="\"$OUT/rayleigh_f2c.c\"" \
      -o "$OUT/librayleigh_f2c.so" "$HERE/ref_harness/rayleigh_f90_harness.c" -lm
  echo "build_ref: built $OUT/librayleigh_f2c.so from $REF/surfmodes/Rayleigh.f90 (startl, SecFunSurf, propup, EinvE, inv2) + bisecim, sort of util.f90 + the C_Interval file"
else
  echo "build_ref: $REF/surfmodes/Love.f90 not present, skipping the translated Love secular function"
fi
# ---- and, where a Fortran compiler exists, the real thing (oracle/build_ref_surfdisp.sh)
if command -v gfortran >/dev/null 2>&1; then "$HERE/build_ref_surfdisp.sh" || true; fi
# ---- stage 1: the reference's own pre-built kd-tree object
if [ ! -f "$REF/utils/libutils.a" ]; then echo "build_ref: $REF/utils/libutils.a not present, skipping"; exit 0; fi
GF="$(python - <<'PY'
import glob, os, sys
try:
    import scipy
    base = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    c = sorted(glob.glob(os.path.join(base, "libgfortran*.so.5*")))
    print(c[0] if c else "")
except Exception:
    print("")
PY
)"
if [ -z "$GF" ]; then echo "build_ref: no libgfortran.so.5 found, skipping"; exit 0; fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
( cd "$TMP" && ar x "$REF/utils/libutils.a" kdtree2.o )
# -no-pie: the object carries R_X86_64_32 relocations (non-PIC build)
gcc -O1 -no-pie -ffp-contract=off -o "$OUT/kdtree2_ref" "$HERE/ref_harness/kdtree2_harness.c" \
    "$TMP/kdtree2.o" "$GF" -Wl,-rpath,"$(dirname "$GF")" -lm
rm -rf "$TMP"
echo "build_ref: built $OUT/kdtree2_ref (libgfortran: $GF)"
