#!/usr/bin/env bash
# Install the UNMODIFIED reference (mynameisfiber/baton) under baseline/_ref (git-ignored, but it
# travels to the GPU box with the gpurun snapshot).  The reference is not a Python package (no
# setup.py / pyproject.toml), so the prescribed `pip install --target` fails; the fallback is a
# verbatim file copy with a checksum manifest.  Nothing in baseline/_ref is edited.
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
DST="$HERE/_ref"
mkdir -p "$DST"
if python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse \
      --target "$DST" "$SRC" >"$DST/pip_install.log" 2>&1; then
  echo "pip install ok" | tee "$DST/INSTALL_STATUS"
else
  echo "pip install failed (reference has no setup.py/pyproject.toml); copying sources verbatim" | tee "$DST/INSTALL_STATUS"
  cp "$SRC"/*.py "$DST"/
  cp "$SRC"/requirements.txt "$SRC"/README.md "$DST"/ 2>/dev/null || true
fi
( cd "$DST" && sha256sum *.py > MANIFEST.sha256 )
( cd "$SRC" && sha256sum *.py ) | diff -q - "$DST/MANIFEST.sha256" >/dev/null && echo "verified: byte-identical to $SRC" | tee -a "$DST/INSTALL_STATUS"
