#!/bin/bash
# round 2, call B: stats-pass breakdown + ncu of the fused scatter kernel, full tests, new bench
mkdir -p gpurun_out
echo "== trace C4"; PLDA_B200_TRACE=1 timeout 300 python scripts/r2_stats_probe.py 5000000 512 50000 1 f32 2>&1 | grep -E "plda_b200 fit|stats_ms" | tail -n 12
echo "== trace C2"; PLDA_B200_TRACE=1 timeout 300 python scripts/r2_stats_probe.py 100000 200 1000 1 f32 2>&1 | grep -E "plda_b200 fit|stats_ms" | tail -n 12
echo "== ncu launch list stats (2M x 512)"; timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_stats_launches.csv python scripts/r2_stats_probe.py 2000000 512 20000 1 f32 > gpurun_out/ncu_stats.log 2>&1; echo "exit=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_stats_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
per=collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[ii], r[ki][:60]), {})[r[mi]] = r[vi]
# last fit only: print the tail
items=list(per.items())[-40:]
for (i,k),m in items:
    print(i, k, m.get('gpu__time_duration.sum'), m.get('dram__bytes_read.sum'), m.get('dram__bytes_write.sum'))
PY
echo "== ncu full scatter kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_syrk -s 2 -c 1 -o gpurun_out/r02_prof_scatter -f python scripts/r2_stats_probe.py 2000000 512 20000 1 f32 > gpurun_out/ncu_scatter.log 2>&1; echo "exit=$?"
echo "== full gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit=$?"; tail -n 12 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit=$?"; tail -n 5 gpurun_out/bench.err; python -c "
import json
j=json.load(open('gpurun_out/bench.json'))
print({k:j[k] for k in ('value','ms_per_step','e2e','e2e_trials','em','cpu_baseline','cpu_best') if k in j})
print(json.dumps(j.get('configs'), indent=1)[:6000])
"
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 | cut -c 1-400
