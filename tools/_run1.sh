set -x
EEGX_NCU_STEP=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches.csv python bench.py --steps 2 --warmup 3 --no-torch-arm > gpurun_out/ncu_step.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/gemm_shapes_in_step.py > gpurun_out/r2_gemm_shapes_in_step.txt 2>&1; echo "shapes rc=$?"
timeout 300 python tools/bench_gemm.py > gpurun_out/r2_bench_gemm_shapes.txt 2>&1; echo "benchgemm rc=$?"
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_sanitizer_memcheck.log python -m pytest tests/test_normalize_gpu.py tests/test_dsp_gpu.py tests/test_fused_gpu.py tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/r2_sanitizer_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r2_sanitizer_memcheck_pytest.log; tail -5 gpurun_out/r2_sanitizer_memcheck.log
