#!/bin/sh
# Places the UNMODIFIED reference package under the git-ignored baseline/_ref/ so that bench.py can time the
# reference's own torch functions (src.quantization / src.cache) beside the kernels on the GPU box, where
# /root/reference does not exist.  `pip install --target baseline/_ref /root/reference` is tried first; the
# reference's build backend (hatchling) is not in this image, so the fallback copies its pure-Python `src`
# package -- exactly the files the wheel would contain.  Nothing under baseline/_ref is tracked by git.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/src" ] || { echo "no reference at $REF"; exit 0; }
mkdir -p "$ROOT/baseline/_ref"
if [ ! -f "$ROOT/baseline/_ref/src/quantization/ops.py" ]; then
    TMP="$(mktemp -d)"; cp -r "$REF" "$TMP/ref"
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target "$ROOT/baseline/_ref" "$TMP/ref" >/dev/null 2>&1 || {
        rm -rf "$ROOT/baseline/_ref/src"; cp -r "$REF/src" "$ROOT/baseline/_ref/src"
        find "$ROOT/baseline/_ref" -name __pycache__ -type d -exec rm -rf {} + 2>/dev/null || true
        echo "pip could not build the reference (no hatchling): copied its src package to baseline/_ref"; }
    rm -rf "$TMP"
fi
