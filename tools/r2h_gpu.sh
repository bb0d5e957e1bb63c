mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
timeout 200 python tools/config_perf.py > gpurun_out/r2h_config_perf.txt 2>&1; cat gpurun_out/r2h_config_perf.txt
for c in "C1q 15" "C3 20" "C1q -5"; do
  set -- $c
  timeout 120 python tools/prof_frontend.py $1 $2 65536 > gpurun_out/r2h_pf_$1_$2.log 2>&1 && \
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_launches_$1_$2.csv \
      python tools/prof_frontend.py $1 $2 32768 > gpurun_out/r2h_ncu_$1.log 2>&1
  cat gpurun_out/r2h_pf_$1_$2.log
done
