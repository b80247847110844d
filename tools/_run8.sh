run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 15 --warmup 4 > gpurun_out/r2_n8_$tag.json 2> gpurun_out/r2_n8_$tag.err; python - <<PY
import json
s=open('gpurun_out/r2_n8_$tag.json').read()
try:
    i=s.index('{"torch_b200"'); d,_=json.JSONDecoder().raw_decode(s[i:])
    print('$tag', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))
except Exception as e: print('$tag FAILED', e)
PY
}
run nch8 NCCL_NVLS_NCHANNELS=8
run nch16 NCCL_NVLS_NCHANNELS=16
run ctas16 NCCL_MAX_CTAS=16
