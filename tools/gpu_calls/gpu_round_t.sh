python bench.py --steps 50 --warmup 5 --krylov-phases > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench exit $?" >> gpurun_out/r02t_bench.err
