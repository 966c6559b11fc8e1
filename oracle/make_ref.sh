#!/usr/bin/env bash
# oracle/make_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Packs the UNMODIFIED reference modules that SURVEY.md section 8a cites (src/dataset, src/metrics, src/losses,
# src/models/{evaluator,probability_helper,tester,trainer,losses,temp_scaling}.py, src/utils) into ONE build
# artefact, oracle/_ref/reference_src.zip, straight from where they lie under /root/reference.  The archive is
# git-ignored (like libslu.so) so no reference source enters the history, but it is NOT gpurun-ignored, so it
# travels to the GPU box, where /root/reference does not exist.  Python imports it with zipimport
# (oracle/ref_arm.py puts "<zip>/src" on sys.path); nothing is unpacked, patched or edited.
#
# Users: bench.py --impl reference / cpu_baseline (the reference's own CPU code timed on the host cores) and
# tests/test_gpu_tester_dropin.py (the reference's Tester.test_epoch driven twice: stock classes vs the
# semanticlidarunc_b200 import swap).  The product package never imports it.
set -euo pipefail
REF="${SLU_REFERENCE_ROOT:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "make_ref: $REF/src not present (GPU box?): keeping whatever $OUT already holds" >&2
    exit 0
fi
mkdir -p "$OUT"
python - "$REF" "$OUT/reference_src.zip" <<'PY'
import hashlib, json, os, sys, zipfile
ref, out = sys.argv[1], sys.argv[2]
keep_dirs = ("dataset", "metrics", "losses", "models", "utils")
names = []
for d in keep_dirs:
    for root, _, files in os.walk(os.path.join(ref, "src", d)):
        for f in sorted(files):
            if f.endswith(".py"):
                names.append(os.path.relpath(os.path.join(root, f), ref))
names.sort()
manifest = {}
tmp = out + ".tmp"
with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
    # explicit directory entries: zipimport needs them to treat src/<pkg> (no __init__.py in the reference) as namespace packages
    for d in sorted({os.path.dirname(n) + "/" for n in names} | {"src/"}):
        z.writestr(zipfile.ZipInfo(d, date_time=(2020, 1, 1, 0, 0, 0)), b"")
    for n in names:
        with open(os.path.join(ref, n), "rb") as fh:
            data = fh.read()
        manifest[n] = hashlib.sha256(data).hexdigest()
        zi = zipfile.ZipInfo(n, date_time=(2020, 1, 1, 0, 0, 0))      # fixed stamps: the archive is reproducible
        zi.compress_type = zipfile.ZIP_DEFLATED
        z.writestr(zi, data)
    zi = zipfile.ZipInfo("MANIFEST.json", date_time=(2020, 1, 1, 0, 0, 0))
    z.writestr(zi, json.dumps({"source": ref, "sha256": manifest}, indent=1, sort_keys=True))
os.replace(tmp, out)
print("make_ref: %d unmodified reference files -> %s" % (len(names), out))
PY
