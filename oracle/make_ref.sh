#!/bin/bash
# TEST / BASELINE INFRASTRUCTURE.  Places the UNMODIFIED reference package
# (johannesulf/nautilus, pure Python) under oracle/_ref/ so that it travels to
# the GPU box with the snapshot (oracle/_ref/ is git-ignored, never committed,
# and not gpurun-ignored).  `bench.py --impl reference` and bench.py's
# cpu_baseline leg import it from there and time the reference's own
# Sampler.add_samples on the host cores; nothing in nautilus_b200/ imports it.
# Run in the build container, where /root/reference exists:
#     bash oracle/make_ref.sh
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC/nautilus" ]; then
  echo "make_ref: $SRC/nautilus not found (nothing to do on the GPU box)"
  exit 0
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$SRC/nautilus" "$HERE/_ref/nautilus"
find "$HERE/_ref" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && find nautilus -name '*.py' -type f | sort | xargs sha256sum ) \
  > "$HERE/_ref/SHA256SUMS"
echo "make_ref: $(wc -l < "$HERE/_ref/SHA256SUMS") files -> $HERE/_ref/nautilus"
