# round 2 evidence session (one B200): full GPU suite, both bench arms, launch list of the bench command, ncu --set full of the
# decoder (source counters included), of the early-exit front end and of the fp16x2 min-sum kernel, parity at 5x scale, smoke
T=${1:-r2k}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --quick --no-c5 > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 100 python tools/prof_decode.py 16384 5 -5 > gpurun_out/${T}_prof_decode.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bp_regular -c 1 -o gpurun_out/${T}_bp_decoder \
    python tools/prof_decode.py 16384 1 -5 > gpurun_out/${T}_ncu_bp.log 2>&1
timeout 100 python tools/prof_frontend.py C1q 15 16384 > gpurun_out/${T}_prof_frontend.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'kmeans_warp|demap_kernel|channel_kernel|encode_kernel' -s 4 -c 4 \
    -o gpurun_out/${T}_frontend python tools/prof_frontend.py C1q 15 16384 > gpurun_out/${T}_ncu_frontend.log 2>&1
KML_ALG=2 timeout 100 python tools/prof_decode.py 16384 5 -5 > gpurun_out/${T}_prof_ms2.log 2>&1 && \
KML_ALG=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ms2_regular -c 1 -o gpurun_out/${T}_ms2 \
    python tools/prof_decode.py 16384 1 -5 > gpurun_out/${T}_ncu_ms2.log 2>&1
KML_PARITY_SCALE=5 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k parity_statistics 2>&1 | grep -E "^peg|^5g|passed|failed" | tee gpurun_out/${T}_parity_at_scale.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
cat gpurun_out/${T}_prof_decode.log gpurun_out/${T}_prof_ms2.log; tail -c 400 gpurun_out/${T}_bench.err
