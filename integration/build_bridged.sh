#!/usr/bin/env bash
# integration/build_bridged.sh -- builds the reference's pyrh C library WITH the B200 bridge compiled in:
#   scratch copy of $REF (default /root/reference, read-only) -> apply integration/pyrh_b200.patch ->
#   compile the source set of setup.py:43-66 with -DPYRH_B200 + integration/pyrh_b200_bridge.c ->
#   oracle/_build/libpyrh_bridged.so, linked against pyrh_b200/csrc/librhb200.so (rpath relative to the library).
# Its rhf1d() has the exact prototype and return struct of rh/rhf1d/pyrh_compute1dray.h:12-36; the per-column work
# runs on the GPU.  The output is test infrastructure (built from the reference's sources): git-ignored, it travels to
# the GPU box like the oracle.  No-op when $REF is absent (prebuilt library is kept).
set -euo pipefail
REF="${REF:-/root/reference}"
here="$(cd "$(dirname "$0")" && pwd)"
root="$(dirname "$here")"
out="$root/oracle/_build"
mkdir -p "$out"
[ -f "$REF/rh/rh.h" ] || { echo "bridge: $REF not present; keeping prebuilt $out/libpyrh_bridged.so"; exit 0; }
[ -f "$root/pyrh_b200/csrc/librhb200.so" ] || { echo "bridge: build librhb200.so first (python -m pyrh_b200.build)"; exit 1; }
W="$(mktemp -d /tmp/pyrh_bridged.XXXXXX)"
trap 'rm -rf "$W"' EXIT
mkdir -p "$W/rh"
cp "$REF/rh/"*.c "$REF/rh/"*.h "$W/rh/"
cp -r "$REF/headers" "$W/headers"
mkdir -p "$W/rh/rhf1d"
cp "$REF/rh/rhf1d/"*.c "$REF/rh/rhf1d/"*.h "$W/rh/rhf1d/"
chmod -R u+w "$W"
( cd "$W" && patch -p1 -s < "$here/pyrh_b200.patch" )

sym="$W/symver.h"          # XDR entry points: glibc compat symbols (no libtirpc in this image), as oracle/build_ref.sh
for s in xdr_bool xdr_double xdr_enum xdr_int xdr_short xdr_string xdr_vector xdrstdio_create; do
  echo "__asm__(\".symver $s,$s@GLIBC_2.2.5\");" >> "$sym"
done
f1d="anglequad feautrier multiatmos formal piecestokes_1D writeflux_xdr bezier_1D hydrostat
     piecewise_1D riiplane pyrh_compute1dray pyrh_solveray project writegeom_xdr
     pyrh_background pyrh_hse pyrh_read_input"
srcs=()
for f in "$W"/rh/*.c; do
  [ "$(basename "$f")" = "collision_Oslo.c" ] && continue
  srcs+=("$f")
done
for n in $f1d; do srcs+=("$W/rh/rhf1d/$n.c"); done
mkdir -p "$W/obj"
CFLAGS="-O2 -fPIC -w -DPYRH_B200 -include $sym -I$W/rh -I$W -I$W/rh/rhf1d -I$root/include"
printf '%s\n' "${srcs[@]}" | xargs -P "$(nproc)" -I{} sh -c \
  'f="{}"; o="'"$W"'/obj/$(echo "$f" | sed "s#'"$W"'/##; s#/#_#g; s#\.c\$#.o#")"; gcc '"$CFLAGS"' -c "$f" -o "$o"'
gcc $CFLAGS -Wall -Wno-unused -c "$here/pyrh_b200_bridge.c" -o "$W/obj/zz_bridge.o"
gcc -shared -o "$out/libpyrh_bridged.so" "$W"/obj/*.o -L"$root/pyrh_b200/csrc" -lrhb200 \
    -Wl,-rpath,'$ORIGIN/../../pyrh_b200/csrc' -lm -lpthread
echo "bridge: built $out/libpyrh_bridged.so"
