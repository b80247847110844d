timeout 600 python -m pytest tests/test_train_gpu.py tests/test_decoder_gpu.py -x -q -m gpu -k "two_gpus or trained" 2>&1 | grep -v Warning | tail -6 | tee gpurun_out/r2_two_gpu_tests.txt
