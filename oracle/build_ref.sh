#!/bin/sh
# Stage the UNMODIFIED reference (a pure-Python tree: nothing to compile) under oracle/_ref/ so that it travels to the
# GPU box with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).  Build container only: the GPU box has no
# /root/reference and uses what was staged here.  Used by: bench.py --impl reference (kind "reference"), the cpu_baseline
# leg of bench.py, tests/test_dropin_reference.py.  TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the product never imports it.
set -e
REF="${ST_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/models" ] || { echo "build_ref: no reference at $REF (nothing staged)"; exit 0; }
rm -rf "$OUT"
mkdir -p "$OUT"
for d in models modules utils trainer; do
    mkdir -p "$OUT/$d"
    for f in "$REF/$d"/*.py; do cp "$f" "$OUT/$d/"; done
done
cp "$REF/train.py" "$REF/translate.py" "$OUT/"
( cd "$REF" && find models modules utils trainer train.py translate.py -maxdepth 1 -name '*.py' | sort | xargs sha256sum ) > "$OUT/SHA256SUMS"
echo "build_ref: staged $(wc -l < "$OUT/SHA256SUMS") files of $REF into $OUT"
