#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference CPU tracker (header-only path of
# hguo/ftk) from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored).
# No reference source is copied into this repository; only a generated config.hh (the file cmake
# would generate from include/ftk/config.hh.in with every FTK_HAVE_* undefined) is written to
# oracle/_ref/gen/ftk/config.hh.   Recipe: SURVEY.md App. B.
set -e
REF=${FTK_REFERENCE_ROOT:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/include/ftk" ]; then
  echo "build_ref.sh: $REF not present; keeping prebuilt oracle/_ref (if any)" >&2
  exit 0
fi
mkdir -p "$OUT/gen/ftk"
sed -e 's/${FTK_VERSION}/0.0.9/' \
    -e 's/#cmakedefine \(.*\) 1/\/* #undef \1 *\//' \
    -e 's/${FTK_FP_PRECISION}/32768/' \
    -e 's/${FTK_CP_MAX_NUM_VARS}/3/' \
    "$REF/include/ftk/config.hh.in" > "$OUT/gen/ftk/config.hh"
SRC="$HERE/ref_harness.cpp"
BIN="$OUT/ftk_ref_oracle"
if [ -x "$BIN" ] && [ "$BIN" -nt "$SRC" ] && [ "$BIN" -nt "$0" ]; then
  exit 0
fi
# -include cmath/limits: include/ftk/numeric/clamp.hh uses std::isnan without <cmath>
# -ffp-contract=off: x86-64 baseline has no FMA contraction; -fwrapv: the reference relies on
# two's-complement wrap of its int64 determinants.
g++ -std=c++17 -O2 -include cmath -include limits -ffp-contract=off -fwrapv -w \
    -I"$OUT/gen" -I"$REF/include" "$SRC" -o "$BIN" -lpthread
echo "built $BIN"
