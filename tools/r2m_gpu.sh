# round 2, after the evidence session: full suite on the final tree, parity at 5x scale (all lines), launch lists of the fused
# path per configuration, final bench line (traffic from this round's capture, e2e through both receiver entry points)
T=${1:-r2m}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.log
KML_PARITY_SCALE=5 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k parity_statistics 2>&1 | grep -E "frames'|passed|failed" | sed 's/^\.*//' | tee gpurun_out/${T}_parity_at_scale.txt
timeout 200 python tools/config_perf.py > gpurun_out/${T}_config_perf.txt 2>&1; cat gpurun_out/${T}_config_perf.txt
for c in "C1q 15" "C4g 15" "C3 20" "C2 10"; do
  set -- $c
  timeout 120 python tools/prof_frontend.py $1 $2 65536 > gpurun_out/${T}_pf_$1.log 2>&1 && \
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_$1.csv \
      python tools/prof_frontend.py $1 $2 32768 > gpurun_out/${T}_ncu_$1.log 2>&1
  cat gpurun_out/${T}_pf_$1.log
done
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
