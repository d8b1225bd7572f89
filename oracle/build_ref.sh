#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the reference's own implementation as a checker.
#
# Compiles /root/reference/mmannot.cpp (read where it lies; nothing is copied into
# the repo) with the reference's Makefile flags (Makefile:7) into oracle/_ref/:
#   mmannot_asis   verbatim source
#   mmannot_fixed  verbatim + the one-line repair of XamRecord::setFlags (mm:606,
#                  unnamed parameter => reads uninitialised `flags`); needed for any
#                  stranded (-s F / -s R) comparison, see SURVEY.md section 0.1
#   mmannot_dump   `fixed` + the interval print at mm:1271 un-commented (feature-order
#                  oracle for the host GTF front-end)
# The patched variants are produced by a sed pipe straight into the compiler; no
# patched source is ever written to disk.
# oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.
set -euo pipefail
REF="${MMANNOT_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
SRC="$REF/mmannot.cpp"
if [ ! -f "$SRC" ]; then
  echo "build_ref.sh: $SRC not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
FLAGS="-DMMSTANDALONE -std=c++11 -pthread -O3 -w"
FIX='s/void setFlags (unsigned int) {/void setFlags (unsigned int f) { flags = f;/'
DUMP='s|^\( *\)//cerr << "\\t" << intervals\[i\] << endl;|\1cerr << "\\t" << intervals[i] << " " << intervals[i].getId() << " " << intervals[i].getStrand() << endl;|'
need() { [ ! -x "$1" ] || [ "$SRC" -nt "$1" ] || [ "$0" -nt "$1" ]; }
if need "$OUT/mmannot_asis";  then g++ "$SRC" $FLAGS -o "$OUT/mmannot_asis" -lz & fi
if need "$OUT/mmannot_fixed"; then sed "$FIX" "$SRC" | g++ -x c++ - $FLAGS -o "$OUT/mmannot_fixed" -lz & fi
if need "$OUT/mmannot_dump";  then sed -e "$FIX" -e "$DUMP" "$SRC" | g++ -x c++ - $FLAGS -o "$OUT/mmannot_dump" -lz & fi
wait
ls -la "$OUT"
