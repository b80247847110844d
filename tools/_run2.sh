timeout 300 python -m pytest tests/test_dsp_gpu.py -x -q -m gpu 2>&1 | tail -1
for v in 0 1; do
EEGX_DSP_VARIANT=$v timeout 120 python bench.py --workload dsp --config long --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('long variant $v', d['ms_per_step'], d['roofline']['frac'])"
done
