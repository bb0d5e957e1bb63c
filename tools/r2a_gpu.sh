# round 2, first GPU call: baseline of the early-exit regime on the round-1 kernels (launch lists + front-end captures)
mkdir -p gpurun_out
python tools/config_perf.py > gpurun_out/r2a_config_perf.txt 2>&1
for c in "C1q 15" "C1p 10" "C3 20" "C2 10" "C4g 15"; do
  set -- $c
  python tools/prof_frontend.py $1 $2 65536 > gpurun_out/r2a_pf_$1.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2a_launches_$1.csv \
      python tools/prof_frontend.py $1 $2 32768 > gpurun_out/r2a_ncu_$1.log 2>&1
done
python tools/prof_frontend.py C1q 15 16384 > gpurun_out/r2a_pf2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'kmeans_warp|demap_kernel' -s 2 -c 2 -o gpurun_out/r2a_frontend_c1 \
    python tools/prof_frontend.py C1q 15 16384 > gpurun_out/r2a_ncu_full.log 2>&1
cat gpurun_out/r2a_config_perf.txt gpurun_out/r2a_pf_*.log
