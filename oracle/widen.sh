#!/usr/bin/env bash
# widen.sh SRC_DIR OUT_DIR -- make a "wide" scratch copy of the reference sources.
#
# Test infrastructure.  The unmodified reference caps a world at 64 dynamic + 64 static
# bodies and 64 manifolds (src/includes/px_body_array.h:11, src/includes/px_manifold_pool.h:8)
# and keeps its sleep bookkeeping in two uint64_t masks (src/px_world.c:143-144).  The
# BASELINE configs need ~256 bodies and ~700 manifolds per world, so the oracle is also built
# from a scratch copy with the capacities raised -- and nothing else changed:
#   * PX_MAX_BODIES 64 -> 255            (255 is the invalid-index sentinel, px_body.h:13)
#   * PX_MAX_MANIFOLDS 64 -> 2048, pool length / loop counter uint8_t -> uint16_t
#   * the "index < 64" asserts dropped   (px_body_array.c:62, px_world.c:158,162)
#   * the two sleep masks become uint64_t[4]; set by body->worldIndex, tested by the sorted
#     position idx -- exactly as the reference does (the index mix-up is preserved)
# The copy lives under build/ (git-ignored); reference sources are never committed.
set -euo pipefail
SRC="$1"
OUT="$2"
rm -rf "$OUT"
mkdir -p "$OUT"
cp -r "$SRC"/. "$OUT"/

inc="$OUT/includes"
n=0
edit() { # edit FILE SED_EXPR EXPECTED_HITS PATTERN
  local f="$1" expr="$2" want="$3" pat="$4"
  local have
  have=$(grep -c -E "$pat" "$f" || true)
  if [ "$have" != "$want" ]; then
    echo "widen.sh: expected $want hits of /$pat/ in $f, found $have" >&2
    exit 1
  fi
  sed -i -E "$expr" "$f"
  n=$((n + have))
}

edit "$inc/px_body_array.h" 's/^#define PX_MAX_BODIES 64$/#define PX_MAX_BODIES 255/' 1 '^#define PX_MAX_BODIES 64$'
edit "$inc/px_manifold_pool.h" 's/^#define PX_MAX_MANIFOLDS 64$/#define PX_MAX_MANIFOLDS 2048/' 1 '^#define PX_MAX_MANIFOLDS 64$'
edit "$inc/px_manifold_pool.h" 's/^  uint8_t length;$/  uint16_t length;/' 1 '^  uint8_t length;$'
edit "$inc/px_manifold_pool.h" 's/for \(uint8_t _i = 0; _i < \(pool\)->length/for (uint16_t _i = 0; _i < (pool)->length/' 1 'for \(uint8_t _i = 0; _i < \(pool\)->length'
edit "$OUT/px_body_array.c" 's/^  assert\(index < 64 .*$/  (void)0;/' 1 '^  assert\(index < 64 '
w="$OUT/px_world.c"
edit "$w" 's/^  uint64_t hasRestingStaticContact = 0;$/  uint64_t hasRestingStaticContact[4] = {0, 0, 0, 0};/' 1 '^  uint64_t hasRestingStaticContact = 0;$'
edit "$w" 's/^  uint64_t shouldWakeUp = 0;$/  uint64_t shouldWakeUp[4] = {0, 0, 0, 0};/' 1 '^  uint64_t shouldWakeUp = 0;$'
edit "$w" 's/^      assert\(body[AB]->worldIndex < 64 .*$/      (void)0;/' 2 '^      assert\(body[AB]->worldIndex < 64 '
edit "$w" 's/(hasRestingStaticContact|shouldWakeUp) \|= \(1ULL << (body[AB]->worldIndex)\);/\1[\2 >> 6] |= (1ULL << (\2 \& 63));/' 4 '(hasRestingStaticContact|shouldWakeUp) \|= \(1ULL << body[AB]->worldIndex\);'
edit "$w" 's/if \((hasRestingStaticContact|shouldWakeUp) & \(1ULL << idx\)\)/if (\1[idx >> 6] \& (1ULL << (idx \& 63)))/' 2 'if \((hasRestingStaticContact|shouldWakeUp) & \(1ULL << idx\)\)'

echo "widen.sh: patched $n lines into $OUT"
