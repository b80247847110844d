timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
