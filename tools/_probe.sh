timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_batched.py -x -q 2>&1 | tail -15 > gpurun_out/n2_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c5 --skip-paged-decode 2> gpurun_out/n2_c5.err | tail -1 > gpurun_out/r2g_bench_n2_c5only.json
