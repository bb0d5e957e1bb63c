# 8-GPU session: multi-GPU tests, bench at N = 8, 4, 2 (weak-scaled headline + the strong-scaled C5 leg through kml_sweep_run)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests -q -m gpu -x -k "multi_gpu or comm_init or pipelined or zero_error" 2>&1 | tail -4 | tee gpurun_out/r2n_pytest.log
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --quick \
     > gpurun_out/r2n_bench_${n}gpu.json 2> gpurun_out/r2n_bench_${n}gpu.err
  tail -c 300 gpurun_out/r2n_bench_${n}gpu.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2n_bench_${n}gpu.json').read().strip().splitlines()[-1])
print($n,'value',round(d['value']),'e2e',round(d['e2e']['value']),'blocking',round(d['e2e']['blocking_call']['value']),'c5',d['c5']['seconds'],d['c5']['frames_per_s'],d['c5']['counters_checksum'])
PY
done
timeout 200 python bench.py --steps 10 --warmup 3 --quick --no-cpu > gpurun_out/r2n_bench_1gpu.json 2> gpurun_out/r2n_bench_1gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2n_bench_1gpu.json').read().strip().splitlines()[-1])
print(1,'value',round(d['value']),'e2e',round(d['e2e']['value']),'blocking',round(d['e2e']['blocking_call']['value']),'f64',round(d['e2e']['reference_types']['value']),'c5',d['c5']['seconds'],d['c5']['frames_per_s'],d['c5']['counters_checksum'], 'traffic', d['roofline']['traffic'])
PY
