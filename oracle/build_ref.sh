#!/usr/bin/env bash
# Installs the UNMODIFIED reference (pure Python) into oracle/_ref/ so that the CPU arm of bench.py
# (`--impl reference`, cpu_baseline.kind = "reference") can time the reference's own implementation on the GPU
# box, where /root/reference does not exist.  oracle/_ref/ is git-ignored (never committed) but travels with the
# repository snapshot.  Test infrastructure: nothing in cosmicfishpie_b200/ imports it.
#   usage: oracle/build_ref.sh [/root/reference]
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -d "$SRC" ] || { echo "reference tree $SRC not found: nothing to do"; exit 0; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$SRC" "$TMP/reference"            # the source tree is read-only; the build writes egg-info
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/reference"
# the reference imports `nautilus` at module level in its likelihood package: the 6-line stub used for the goldens
cp -r "$HERE/../tools/_stubs/nautilus.py" "$HERE/_ref/nautilus.py"
echo "installed: $(ls "$HERE/_ref" | tr '\n' ' ')"
