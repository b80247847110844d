timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_train_gpu.py -x -q -m gpu 2>&1 | grep -v Warning | tail -12
