mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests -q -m gpu -x -k "multi_gpu or comm_init or debug_mode or kmeans_channel or two_streams or histogram" 2>&1 | tail -8 > gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --quick > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err
tail -c 600 gpurun_out/r2g_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_2gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'c5',json.dumps(d.get('c5'))[:700])
PY
timeout 120 python bench.py --steps 3 --warmup 3 --quick --no-cpu > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_1gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'kmeans',d['kmeans'],'c5',json.dumps(d.get('c5'))[:500])
PY
