#!/bin/bash
# Builds the drop-in evidence under build/compat/:
#   dropin_check.out            tests/cpp/dropin_check.cu (ours) against include/ + libgnnagg.so
#   Figure9_main.out, ...       the reference's UNMODIFIED drivers, compiled from where they lie under
#                               /root/reference against OUR include/ (only when that tree exists)
# Binaries are git-ignored (build/) but travel to the GPU box with gpurun.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/build/compat"
mkdir -p "$OUT"
CCBIN=""; [ -x /usr/bin/g++ ] && CCBIN="-ccbin /usr/bin/g++"
FLAGS="$CCBIN -std=c++17 -O2 -w -gencode arch=compute_100a,code=sm_100a -I$ROOT/include -L$ROOT/gnn-computing_b200/lib -lgnnagg -lcurand -lcublas -Xlinker -rpath,$ROOT/gnn-computing_b200/lib -Xlinker -rpath,\$ORIGIN/../../gnn-computing_b200/lib"
build() { # src out
  local stale=0
  for h in "$ROOT"/include/*.h; do [ "$h" -nt "$2" ] && stale=1; done
  if [ ! -f "$2" ] || [ "$1" -nt "$2" ] || [ $stale = 1 ]; then
    nvcc $FLAGS "$1" -o "$2"
  fi
}
build "$ROOT/tests/cpp/dropin_check.cu" "$OUT/dropin_check.out"
build "$ROOT/tests/cpp/dist_check.cu" "$OUT/dist_check.out"
REF="${REF:-/root/reference}"
if [ -d "$REF/Figure9" ]; then
  build "$REF/Figure9/main.cu"    "$OUT/Figure9_main.out"
  build "$REF/Figure10/main_a.cu" "$OUT/Figure10_main_a.out"
  build "$REF/Figure10/main_b.cu" "$OUT/Figure10_main_b.out"
fi
echo "compat binaries in $OUT"
