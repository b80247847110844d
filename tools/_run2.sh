EEGX_DSP_VARIANT=7 timeout 300 python -m pytest tests/test_dsp_gpu.py -x -q -m gpu 2>&1 | tail -5
EEGX_DSP_PROF=1 EEGX_DSP_VARIANT=7 timeout 120 python bench.py --workload dsp --steps 3 --warmup 3 2>&1 | grep "prof" | tail -2
EEGX_DSP_VARIANT=7 timeout 120 python bench.py --workload dsp --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])"
