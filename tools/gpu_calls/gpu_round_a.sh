python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?" >> gpurun_out/r02a_bench.err
python tools/exp_pt2_cap.py > gpurun_out/r02a_pt2cap.log 2>&1
