# 1-GPU session on the final tree: full GPU suite, smoke, the default bench (both arms), launch list of the bench command
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2q_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2q_smoke.log
timeout 600 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 400 gpurun_out/r2q_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',d['roofline']['frac'],'c5',d['c5']['seconds'],d['c5']['counters_checksum'])
for c in d.get('configs',[]): print(c['workload'], round(c['value']), c['roofline']['frac'], c.get('throughput_mode'))
r=json.loads(open('gpurun_out/r2q_bench_reference.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('cpu_baseline'))
PY
