#!/bin/bash
# Copies the UNMODIFIED reference packages of the hot path (sloter/, timm/) from /root/reference into the git-ignored
# baseline/_ref/ (it travels to the GPU box with gpurun, it never enters the history), so that `bench.py --impl reference`
# and bench.py's `gpu_eager_baseline` leg run the reference's own SlotModel.forward there.  Nothing is edited: the two
# import shims the 2020 code needs on torch 2.x live in baseline/refload.py.
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
if [ ! -f "$SRC/sloter/slot_model.py" ]; then echo "fetch_ref: no reference at $SRC (nothing copied)"; exit 0; fi
mkdir -p "$DST"
rm -rf "$DST/sloter" "$DST/timm"
cp -r "$SRC/sloter" "$SRC/timm" "$DST/"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
(cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown) > "$DST/REFERENCE_COMMIT"
echo "fetch_ref: $(find "$DST" -name '*.py' | wc -l) reference files in $DST"
