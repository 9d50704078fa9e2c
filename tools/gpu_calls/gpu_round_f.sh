# N = 2: multi-GPU parity (tests/test_gpu_multi.py world 2) and the bench with its parity object
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/r02f_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "bench exit $?" >> gpurun_out/r02f_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r02f_ref_n2.json 2> gpurun_out/r02f_ref_n2.err
