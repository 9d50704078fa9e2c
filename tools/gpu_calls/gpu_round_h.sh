# N = 2: multi-GPU parity + bench (after the Davidson / PT2 planning / packed-build changes)
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/r02h_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 3 --krylov-phases --no-cpu-baseline > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; echo "bench exit $?" >> gpurun_out/r02h_bench_n2.err
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed or pt2 or peer or skqd" > gpurun_out/r02h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_pytest.log
