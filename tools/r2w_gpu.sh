# evidence on the final tree (after the fp16x2 plan kernel, the 64QAM grid demapper, metric-decode reuse, four simulate lanes): full GPU suite, bench (both arms), per-config fused-path
# throughput, launch list of C3
T=${1:-r2w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/${T}_pytest.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
timeout 200 python tools/config_perf.py > gpurun_out/${T}_config_perf.txt 2>&1; cat gpurun_out/${T}_config_perf.txt
timeout 120 python tools/prof_frontend.py C3 20 65536 > gpurun_out/${T}_pf_C3.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_C3.csv \
    python tools/prof_frontend.py C3 20 32768 > gpurun_out/${T}_ncu_C3.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',d['roofline']['frac'],'c5',d['c5']['seconds'],d['c5']['counters_checksum'])
for c in d.get('configs',[]): print(c['workload'], round(c['value']), c['roofline']['frac'], c.get('throughput_mode'))
PY
