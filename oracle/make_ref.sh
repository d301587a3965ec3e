#!/bin/sh
# oracle/make_ref.sh -- stage the UNMODIFIED reference modules of the hot path under
# oracle/_ref/ so that `bench.py --impl reference` and bench.py's cpu_baseline can time the reference's own
# PyTorch implementation on the GPU box's host cores (kind: "reference").
#
# TEST / MEASUREMENT INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (no reference source enters the
# history) but not gpurun-ignored (it travels to the GPU box like a built .so).  The reference is pure
# Python: "building" it is staging the files where they can be imported; nothing is edited.
# Run from anywhere:  sh oracle/make_ref.sh [/path/to/reference]
set -e
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/backbones" ] || { echo "make_ref: $REF is not the reference tree" >&2; exit 1; }
rm -rf "$OUT"
mkdir -p "$OUT/backbones/ops/modules" "$OUT/dinov2/layers" "$OUT/dinov2/models" "$OUT/segloss"
# a1-a3  backbones/ops/modules/ms_deform_attn.py      a4-a6  backbones/adapter_blocks.py
# f1     backbones/encoders.py                         f2     backbones/decoders.py (+ unet_parts.py, its import)
# a7-a10 dinov2/layers/*.py                            a11    dinov2/models/*.py      loss   segloss/dice.py
cp "$REF"/backbones/ops/modules/__init__.py "$REF"/backbones/ops/modules/ms_deform_attn.py "$OUT/backbones/ops/modules/"
cp "$REF"/backbones/adapter_blocks.py "$REF"/backbones/encoders.py "$REF"/backbones/decoders.py \
   "$REF"/backbones/unet_parts.py "$OUT/backbones/"
cp "$REF"/dinov2/__init__.py "$OUT/dinov2/"
cp "$REF"/dinov2/layers/*.py "$OUT/dinov2/layers/"
cp "$REF"/dinov2/models/*.py "$OUT/dinov2/models/"
cp "$REF"/segloss/dice.py "$OUT/segloss/"
( cd "$REF" && { git rev-parse HEAD 2>/dev/null || echo unknown; } ) > "$OUT/REVISION"
find "$OUT" -name '*.py' | sort | xargs sha256sum > "$OUT/SHA256SUMS"
echo "make_ref: staged $(find "$OUT" -name '*.py' | wc -l) reference files in $OUT"
