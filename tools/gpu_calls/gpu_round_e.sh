python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
python bench.py --steps 20 --warmup 3 --pt2-c4-sources 0 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench exit $?" >> gpurun_out/r02e_bench.err
