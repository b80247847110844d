run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 15 --warmup 4 > gpurun_out/r2_n8_$tag.json 2> gpurun_out/r2_n8_$tag.err; echo "$tag rc=$?"; python - <<PY
import json
for l in open('gpurun_out/r2_n8_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))
PY
}
run default X=1
run maxctas8 NCCL_MAX_CTAS=8
run nvls NCCL_ALGO=NVLS
grep -h "NVLS\|error\|Error" gpurun_out/r2_n8_nvls.err | head -5
