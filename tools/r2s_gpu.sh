# final evidence session of round 2 (one B200): full GPU suite, parity at 5x scale, both bench arms, launch list of the bench
# command, ncu --set full of the decoder (source counters) and of the layered min-sum kernel, smoke
T=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.log
KML_PARITY_SCALE=5 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k parity_statistics 2>&1 | grep -E "frames'|passed|failed" | sed 's/^\.*//' | tee gpurun_out/${T}_parity_at_scale.txt
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --quick --no-c5 > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 100 python tools/prof_decode.py 16384 5 -5 > gpurun_out/${T}_prof_decode.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bp_regular -c 1 -o gpurun_out/${T}_bp_decoder \
    python tools/prof_decode.py 16384 1 -5 > gpurun_out/${T}_ncu_bp.log 2>&1
KML_ALG=3 timeout 100 python tools/prof_decode.py 8192 5 -5 5GLDPCBG2a3_R12_K960.txt 4bit_16QAM_Gray.txt > gpurun_out/${T}_prof_layered.log 2>&1 && \
KML_ALG=3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ms_layered -c 1 -o gpurun_out/${T}_layered \
    python tools/prof_decode.py 8192 1 -5 5GLDPCBG2a3_R12_K960.txt 4bit_16QAM_Gray.txt > gpurun_out/${T}_ncu_layered.log 2>&1
KML_ALG=1 timeout 100 python tools/prof_decode.py 8192 5 -5 5GLDPCBG2a3_R12_K960.txt 4bit_16QAM_Gray.txt > gpurun_out/${T}_prof_ms5g.log 2>&1
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
cat gpurun_out/${T}_prof_decode.log gpurun_out/${T}_prof_layered.log gpurun_out/${T}_prof_ms5g.log; tail -c 300 gpurun_out/${T}_bench.err
