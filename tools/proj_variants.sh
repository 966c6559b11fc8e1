#!/usr/bin/env bash
# scratch: projection kernel variants (points per thread x CTAs per SM), 16 HDL-64 / OS1-128 scans, graph-replay timing
for lib in libslu.so libslu_pb1.so libslu_pb2.so; do
  for cap in 8 16 32; do
    echo "== $lib cap=$cap"
    SLU_LIB_PATH=$PWD/semanticlidarunc_b200/$lib SLU_PT_CTAS_PER_SM=$cap python tools/kernel_report.py "projection hdl64 B=16" "projection os1-128 B=16" "back-projection hdl64 B=16" 2>/dev/null | python -c "import json,sys; [print(' ', r['stage'], r['ms']) for r in json.load(sys.stdin)['rows']]"
  done
done
