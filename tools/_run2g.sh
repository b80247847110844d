run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 15 --warmup 4 > gpurun_out/r2_n2_$tag.json 2> gpurun_out/r2_n2_$tag.err; echo "rc=$?"; python - <<PY
import json
ok=False
for l in open('gpurun_out/r2_n2_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value'])); ok=True
if not ok: print('$tag FAILED')
PY
grep -h -i "warn\|error\|NVLS\|registr" gpurun_out/r2_n2_$tag.err | head -5
}
run reg EEGX_NCCL_REGISTER=1 NCCL_DEBUG=WARN
run noreg EEGX_NCCL_REGISTER=0
