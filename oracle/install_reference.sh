#!/bin/sh
# TEST / BENCH INFRASTRUCTURE -- installs the UNMODIFIED reference into baseline/_ref (git-ignored, but it travels
# to the GPU box with the repository snapshot), so that `bench.py --impl reference` and `cpu_baseline` can time the
# reference's own classes there.  /root/reference itself does not exist on the GPU box.
#
# The reference's 2017 versioneer.py calls configparser.SafeConfigParser / readfp, which Python 3.12 removed; the
# two names are supplied from OUTSIDE the reference tree through a sitecustomize on PYTHONPATH (the sources are
# not touched; `cmp` below proves the installed files are byte-identical to /root/reference).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"          # the build writes into the source tree; /root/reference is read-only
mkdir -p "$TMP/shim"
cat > "$TMP/shim/sitecustomize.py" <<'PY'
import configparser
if not hasattr(configparser, "SafeConfigParser"):
    configparser.SafeConfigParser = configparser.ConfigParser
if not hasattr(configparser.ConfigParser, "readfp"):
    configparser.ConfigParser.readfp = lambda self, fp, filename=None: self.read_file(fp, filename)
PY
rm -rf "$ROOT/baseline/_ref"
PYTHONPATH="$TMP/shim" python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target "$ROOT/baseline/_ref" "$TMP/ref"
for f in __init__.py dispersion.py optimize.py stencil.py weight.py; do
    cmp "$ROOT/baseline/_ref/extended_stencil/$f" "$REF/extended_stencil/$f"
done
rm -rf "$TMP"
echo "installed the unmodified reference into $ROOT/baseline/_ref"
