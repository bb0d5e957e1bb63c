mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r2p_pytest.log
for c in "C1q 15" "C1q 10" "C1p 10" "C4g 15"; do set -- $c; timeout 100 python tools/prof_frontend.py $1 $2 131072; done
timeout 100 python tools/prof_frontend.py C1q 15 16384 > /dev/null 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2p_launches_C1q.csv python tools/prof_frontend.py C1q 15 32768 > gpurun_out/r2p_ncu.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --quick --no-cpu > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'blocking',round(d['e2e']['blocking_call']['value']),'c5',d['c5']['seconds'],d['c5']['frames_per_s'],d['c5']['counters_checksum'])
PY
