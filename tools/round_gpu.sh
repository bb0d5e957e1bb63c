# One GPU session for the round's evidence: full GPU test suite, bench (both arms), ncu launch list of the bench command,
# ncu full capture of the dominant kernel.  Everything lands in gpurun_out/ with the given tag.
TAG=${1:-r1h}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --quick > gpurun_out/ncu_launch_$TAG.log 2>&1
python tools/prof_decode.py 16384 5 -5 > gpurun_out/prof_$TAG.log 2>&1 && \
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:bp_regular -c 1 -o gpurun_out/bp_$TAG \
    python tools/prof_decode.py 16384 1 -5 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -c 600 gpurun_out/bench_$TAG.json; cat gpurun_out/prof_$TAG.log
