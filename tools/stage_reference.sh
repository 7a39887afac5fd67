#!/bin/bash
# Stage the UNMODIFIED reference package under the git-ignored baseline/_ref/ (it travels to the GPU box with the
# snapshot, where /root/reference does not exist) so that `bench.py --impl reference` can time the reference's own
# code (cpu_baseline.kind = "reference").  The reference's build backend is poetry-core, which is not in this image
# and cannot be fetched (no network): `pip install /root/reference` fails with "No module named 'poetry'".  The
# package is pure Python, so the same files are installed from a copy under /tmp whose ONLY change is the
# [build-system] table (setuptools instead of poetry-core); no file of the package itself is touched.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${1:-/root/reference}
TMP=$(mktemp -d /tmp/refcopy.XXXXXX)
cp -r "$SRC/pyhalomodel" "$TMP/pyhalomodel"
cat > "$TMP/pyproject.toml" <<'EOT'
[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"
[project]
name = "pyhalomodel"
version = "1.0.2"
[tool.setuptools]
packages = ["pyhalomodel"]
EOT
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP" 2>&1 | tail -2
rm -rf "$TMP"
python - <<EOT
import sys
sys.path.insert(0, "$ROOT/baseline/_ref")
import pyhalomodel, os
assert os.path.dirname(pyhalomodel.__file__).startswith("$ROOT/baseline/_ref"), pyhalomodel.__file__
print("staged", pyhalomodel.__file__)
EOT
