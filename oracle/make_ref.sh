#!/bin/sh
# oracle/make_ref.sh -- stage the UNMODIFIED reference step loop under oracle/_ref/ so that the CPU
# legs of bench.py can time the reference itself (kind "reference") instead of the restatement in
# oracle/py_loop.py. Run in the builder container, where /root/reference exists; oracle/_ref/ is
# git-ignored (never part of the history) but travels to the GPU box with the snapshot, like the
# built .so files. Only the pure-Python files of the step path are staged: the engine package and
# the wrapper module (numpy + stdlib only). Nothing in the product imports it
# (tests/test_capi_surface.py::test_product_never_imports_the_oracle).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
if [ ! -d "$REF/inversus" ] || [ ! -f "$REF/inversus_rl/env_wrappers.py" ]; then
    echo "make_ref: no reference at $REF (nothing staged)"; exit 0
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref/inversus" "$HERE/_ref/inversus_rl"
for f in __init__.py core.py game_types.py config.py; do cp "$REF/inversus/$f" "$HERE/_ref/inversus/$f"; done
cp "$REF/inversus_rl/env_wrappers.py" "$HERE/_ref/inversus_rl/env_wrappers.py"
# the reference's own inversus_rl/__init__.py pulls in torch-based modules that are not on this path
: > "$HERE/_ref/inversus_rl/__init__.py"
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$HERE/_ref/SOURCE_COMMIT"
echo "make_ref: staged $(find "$HERE/_ref" -name '*.py' | wc -l) files from $REF into $HERE/_ref"
