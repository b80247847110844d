timeout 300 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "two_gpus" 2>&1 | tail -1
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 15 --warmup 4 > gpurun_out/r2_n2_$tag.json 2> gpurun_out/r2_n2_$tag.err; python - <<PY
import json
s=open('gpurun_out/r2_n2_$tag.json').read()
try:
    i=s.index('{"torch_b200"'); d,_=json.JSONDecoder().raw_decode(s[i:])
    print('$tag', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))
except Exception as e: print('$tag FAILED', e)
PY
}
run sidejoin X=1
run mainjoin EEGX_BOUNDARY_JOIN=1
run sidejoin2 X=1
