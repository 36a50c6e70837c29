#!/bin/bash
# quick timings of the biallelic sweep kernel at config 4 for a few grid shapes, then one ncu capture (run under gpurun)
mkdir -p gpurun_out
tag=${1:-snp}
B="python bench.py --workload c4 --steps 20 --warmup 3 --no-e2e --no-cpu"
for nb in "" 6 12 18 27; do
  IG_SNP_NBLK=$nb $B 2> gpurun_out/${tag}_nb$nb.err | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    r = d['roofline']; print('nblk=$nb', d['config']['geometry']['nblk'], 'ms/step', round(d['ms_per_step'],4), 'zq_ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4))
" | tee -a gpurun_out/${tag}_times.txt
done
$B --steps 3 > gpurun_out/${tag}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:zq_snp -s 4 -c 2 -f -o gpurun_out/${tag}_prof $B --steps 3 > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}_prof.ncu-rep
