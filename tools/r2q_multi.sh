# N-GPU session on the final tree: multi-GPU tests, bench at N = 8, 4, 2, 1 (weak-scaled headline + strong-scaled C5 leg via kml_sweep_run)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests -q -m gpu -x -k "multi_gpu or comm_init" 2>&1 | tail -4 | tee gpurun_out/r2q_multi_pytest.log
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 \
     > gpurun_out/r2q_bench_${n}gpu.json 2> gpurun_out/r2q_bench_${n}gpu.err
  tail -c 300 gpurun_out/r2q_bench_${n}gpu.err
done
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench_1gpu.json 2> gpurun_out/r2q_bench_1gpu.err
for n in 8 4 2 1; do python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_bench_${n}gpu.json').read().strip().splitlines()[-1])
print($n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'c5',d['c5']['seconds'],d['c5']['frames_per_s'],d['c5']['counters_checksum'], d['c5'].get('timing'))
PY
done
