#!/usr/bin/env bash
# oracle/build_ref.sh <scalar|simd> -- TEST INFRASTRUCTURE ONLY.
# Compiles the unmodified reference C sources *where they lie* under $REF
# (default /root/reference, read-only) plus our probe.c into
# oracle/_ref/liboracle_<variant>.so.  Source set = reference setup.py:43-66.
# Flags mirror the reference's distutils build (-O2 -fPIC, no -march, no
# fast-math; -DSIMDON on x86 per setup.py:39-41 for the "simd" variant).
set -euo pipefail
variant="${1:-scalar}"
REF="${REF:-/root/reference}"
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/_ref"
obj="$out/obj_$variant"
mkdir -p "$obj"
[ -f "$REF/rh/rh.h" ] || { echo "oracle: $REF not present; keeping prebuilt $out"; exit 0; }

defs=""
[ "$variant" = "simd" ] && defs="-DSIMDON"

# XDR entry points: bind to glibc's compat symbols (no libtirpc in this image)
sym="$out/symver.h"
: > "$sym"
for s in xdr_bool xdr_double xdr_enum xdr_int xdr_short xdr_string xdr_vector xdrstdio_create; do
  echo "__asm__(\".symver $s,$s@GLIBC_2.2.5\");" >> "$sym"
done

f1d="anglequad feautrier multiatmos formal piecestokes_1D writeflux_xdr bezier_1D hydrostat
     piecewise_1D riiplane pyrh_compute1dray pyrh_solveray project writegeom_xdr
     pyrh_background pyrh_hse pyrh_read_input"
srcs=()
for f in "$REF"/rh/*.c; do
  [ "$(basename "$f")" = "collision_Oslo.c" ] && continue
  srcs+=("$f")
done
for n in $f1d; do srcs+=("$REF/rh/rhf1d/$n.c"); done

CFLAGS="-O2 -fPIC -w $defs -include $sym -I$REF/rh -I$REF"
printf '%s\n' "${srcs[@]}" | xargs -P "$(nproc)" -I{} sh -c \
  'f="{}"; o="'"$obj"'/$(echo "$f" | sed "s#'"$REF"'/##; s#/#_#g; s#\.c\$#.o#")"; gcc '"$CFLAGS"' -c "$f" -o "$o"'
gcc -O2 -fPIC -w $defs -I"$REF/rh" -I"$REF" -c "$here/probe.c" -o "$obj/zz_probe.o"

wrap=""
for s in rlk_opacity writeBackground Piece_Stokes_Bezier3_1D Piecewise_Bezier3_1D Feautrier \
         Formal Opacity addtoGamma addtoRates statEquil Accelerate Iterate updatePopulations SolveLinearEq solveSpectrum \
         Piecewise_1D Piecewise_Linear_1D Piece_Stokes_1D MolecularOpacity passive_bb \
         Thomson Hminus_bf Hminus_ff OH_bf_opac CH_bf_opac Hydrogen_bf Hydrogen_ff Rayleigh H2plus_ff Rayleigh_H2 H2minus_ff Metal_bf ChemicalEquilibrium; do
  wrap="$wrap -Wl,--wrap=$s"
done
gcc -shared -o "$out/liboracle_$variant.so" "$obj"/*.o $wrap -lm -lpthread
echo "oracle: built $out/liboracle_$variant.so"
