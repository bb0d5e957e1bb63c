mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "minsum or peg8064 or demapper or resolver" 2>&1 | tail -8 > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
timeout 120 python tools/prof_frontend.py C3 20 65536 > gpurun_out/r2i_pf_C3.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2i_launches_C3.csv \
    python tools/prof_frontend.py C3 20 32768 > gpurun_out/r2i_ncu_C3.log 2>&1
cat gpurun_out/r2i_pf_C3.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-c5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'kmeans',d['kmeans']); print(json.dumps(d['throughput_mode'],indent=0)[:900]); print(d['early_exit_15dB'])
print([ (c['workload'], round(c['value']), round(c['roofline']['frac'],3)) for c in d.get('configs',[])])
PY
