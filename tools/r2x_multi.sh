# last multi-GPU check on the final tree: multi-GPU tests, bench under torchrun at N = 8 and 2 (C5 leg through kml_sweep_run)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "multi_gpu or comm_init" 2>&1 | tail -3 | tee gpurun_out/r2x_multi_pytest.log
for n in 8 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 \
     > gpurun_out/r2x_bench_${n}gpu.json 2> gpurun_out/r2x_bench_${n}gpu.err
  tail -c 200 gpurun_out/r2x_bench_${n}gpu.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2x_bench_${n}gpu.json').read().strip().splitlines()[-1])
print($n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'c5',d['c5']['seconds'],d['c5']['frames_per_s'],d['c5']['counters_checksum'])
PY
done
