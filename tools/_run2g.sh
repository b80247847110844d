run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 15 --warmup 4 > gpurun_out/r2_n2_$tag.json 2> gpurun_out/r2_n2_$tag.err; python - <<PY
import json
ok=False
for l in open('gpurun_out/r2_n2_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value'])); ok=True
if not ok: print('$tag FAILED')
PY
}
run default X=1
run hiprio TORCH_NCCL_HIGH_PRIORITY=1
run ctas16 NCCL_MAX_CTAS=16
run ctas32 NCCL_MAX_CTAS=32
run nooverlap EEGX_OVERLAP_ALLREDUCE=0
