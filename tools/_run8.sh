timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --config long --steps 15 --warmup 4 > gpurun_out/r2_bench_long_n8.json 2> gpurun_out/r2_bench_long_n8.err; echo "rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r2_bench_long_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('long n8', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['dsp']['value'], d['config']['workload'][:40])
PY
tail -3 gpurun_out/r2_bench_long_n8.err
