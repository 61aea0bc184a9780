timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_paged_quant.py -x -q 2>&1 | tail -15 > gpurun_out/n2_tests.log
