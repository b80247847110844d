for v in 11 12 13; do
EEGX_DSP_VARIANT=$v timeout 300 python -m pytest tests/test_dsp_gpu.py -x -q -m gpu 2>&1 | tail -1
EEGX_DSP_VARIANT=$v timeout 120 python bench.py --workload dsp --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', d['ms_per_step'], d['roofline']['frac'])"
done
timeout 120 python bench.py --workload dsp --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default', d['ms_per_step'], d['roofline']['frac'])"
