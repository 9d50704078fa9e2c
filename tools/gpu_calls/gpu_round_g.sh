# N = 8: multi-GPU parity at world 2/4/8 and the bench with parity object + Davidson phases
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/r02g_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02g_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 50 --warmup 5 --krylov-phases > gpurun_out/r02g_bench_n8.json 2> gpurun_out/r02g_bench_n8.err; echo "bench exit $?" >> gpurun_out/r02g_bench_n8.err
