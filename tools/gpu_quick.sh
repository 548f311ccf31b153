#!/usr/bin/env bash
# Quick GPU pass: parity tests + bench (no ncu). Usage: tools/gpu_quick.sh <tag> [pytest -k expr]
TAG="${1:-q}"; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.log 2>&1; echo "bench rc=$?" >> $OUT/bench_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log; python - <<PY
import json
for l in open("$OUT/bench_$TAG.log"):
    if l.startswith("{"):
        d=json.loads(l); print({k:d[k] for k in ("value","per_frame")}, d["e2e"]["value"], d["roofline"]["frac"])
PY
