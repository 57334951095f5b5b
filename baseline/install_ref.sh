#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with the
# gpurun snapshot) so that bench.py can time the reference's own Python path on the bench box's
# host cores.  Run in the build container, where /root/reference exists.
#
# The reference's setup.py lists its packages by hand and forgets two directories its own factory
# imports at start-up (helpers/factory.py:106 imports ai_safety_gridworlds.environments.aintelope;
# the experiments/ overlays are found the same way), so a plain pip install cannot even construct an
# environment.  The two missing package directories are added from the same source tree, unmodified.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"                              # /root/reference is read-only; the build writes into its tree
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target "$ROOT/baseline/_ref" "$TMP/ref" >/dev/null
for d in ai_safety_gridworlds/environments/aintelope ai_safety_gridworlds/experiments; do
  if [ -d "$REF/$d" ] && [ ! -d "$ROOT/baseline/_ref/$d" ]; then cp -r "$REF/$d" "$ROOT/baseline/_ref/$d"; fi
done
find "$ROOT/baseline/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
rm -rf "$TMP"
echo "installed: $(ls "$ROOT/baseline/_ref")"
