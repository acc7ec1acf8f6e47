#!/usr/bin/env bash
# Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, copied from where they lie under
# /root/reference into the git-ignored (but gpurun-shipped) directory oracle/_ref/, so that the GPU box -- where
# /root/reference does not exist -- can run the real reference: `bench.py --impl reference` (kind "reference"),
# the `reference`-marked tests and the executed drop-in test.  Nothing under oracle/_ref/ is ever committed
# (.gitignore) and the product path (eagcn_b200/) never imports it (tests/test_abi.py enforces that).
#   layers.py  the layer stack (hot path)            models.py  EAGCN (caller)
#   utils.py   weights_init / weight_tensor / collate neural_fp.py  imported by utils.py at module top
set -euo pipefail
SRC="${EAGCN_REFERENCE_SRC:-/root/reference/eagcn_pytorch}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$HERE/oracle/_ref"
if [ ! -f "$SRC/layers.py" ]; then
  echo "[make_oracle_ref] $SRC not present: keeping whatever is in $DST" >&2
  exit 0
fi
mkdir -p "$DST"
for f in layers.py models.py utils.py neural_fp.py; do
  cp -f "$SRC/$f" "$DST/$f"
done
( cd "$SRC" && sha256sum layers.py models.py utils.py neural_fp.py ) > "$DST/SHA256SUMS"
echo "[make_oracle_ref] copied 4 reference modules into $DST"
