#!/bin/bash
# tools/kernel_lab.sh -- time experimental builds of the sweep kernel on one B200 (run under gpurun).
# Each variant is a full library built by `make -C instruct_b200/csrc BUILD=build/v_<name> OUT=../variants/libig_<name>.so
# EXTRA="-DIG_FAST_BUILD <flags>"`; bench.py loads it through IG_LIB.  One JSON line per variant in gpurun_out/lab_<tag>.jsonl.
tag=${1:-lab}; shift
mkdir -p gpurun_out
out=gpurun_out/lab_$tag.jsonl
: > $out
for v in "$@"; do
  IG_LIB=$PWD/instruct_b200/variants/libig_$v.so timeout 300 python bench.py --workload c4 --steps 30 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/lab_${tag}_$v.err | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    r = d['roofline']
    print(json.dumps({'variant': '$v', 'ms_per_step': d['ms_per_step'], 'zq_ms': r['avg_launch_ms'], 'frac': r['frac'], 'clocks': d.get('clocks')}))
" >> $out
  tail -1 $out
done
