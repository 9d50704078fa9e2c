#!/bin/sh
# Vendor the UNMODIFIED reference (pure Python) into oracle/_ref/ so that it travels to the GPU
# box with the gpurun snapshot: oracle/_ref/ is git-ignored (never committed) but not
# gpurun-ignored.  Used only as a CHECKER / CPU baseline: tests/test_gpu_pipeline.py runs the
# reference's own pipeline (a) unpatched on the CPU and (b) with the engine's classes patched in
# (INTEGRATION.md section 3), and bench.py --impl reference times its get_connections.
# Nothing under flow_guided_krylov_b200/ imports it.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC/src" ]; then
    echo "vendor_ref: $SRC/src not found (nothing vendored)"; exit 0
fi
[ -d "$ROOT/oracle/_ref/src" ] && chmod -R u+w "$ROOT/oracle/_ref/src"
rm -rf "$ROOT/oracle/_ref/src"
mkdir -p "$ROOT/oracle/_ref"
cp -r "$SRC/src" "$ROOT/oracle/_ref/src"
chmod -R u+w "$ROOT/oracle/_ref/src"
find "$ROOT/oracle/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "vendor_ref: copied $SRC/src -> oracle/_ref/src ($(find "$ROOT/oracle/_ref/src" -name '*.py' | wc -l) files)"
