#!/bin/sh
# Vendors the UNMODIFIED reference into the git-ignored baseline/_ref/ so that `bench.py --impl reference_gpu` can run
# the reference's own PyTorch model classes on the GPU box (where /root/reference does not exist).
# 1. the sanctioned offline install (the build writes into its source tree, hence the /tmp copy; --no-deps because
#    timm / diffusers / deepspeed / ... are not in the wheelhouse);
# 2. upstream declares `packages = ["dexbotic"]` (pyproject.toml:100) and is used as an editable install, so the wheel
#    carries only dexbotic/{client,constants}.py: the sub-packages are completed from the same source tree.
# Nothing under baseline/_ref/ is tracked by git or imported by the product path.
set -e
cd "$(dirname "$0")/.."
REF=${1:-/root/reference}
rm -rf baseline/_ref /tmp/_refcopy
cp -r "$REF" /tmp/_refcopy
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/_refcopy
cp -rn "$REF"/dexbotic/. baseline/_ref/dexbotic/
find baseline/_ref -name __pycache__ -type d -prune -exec rm -rf {} +
rm -rf /tmp/_refcopy
du -sh baseline/_ref
