timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_gpu_quant.py tests/test_gpu_cache.py tests/test_gpu_paged_quant.py -x -q 2>&1 | tail -3 > gpurun_out/final_tests.log
timeout 600 python bench.py --workload c3 --skip-e2e --skip-cpu > gpurun_out/r2g_bench_n1_c3only.json 2> gpurun_out/r2g.err
echo rc=$? >> gpurun_out/final_tests.log
