python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02v_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02v_pytest.log
python bench.py --steps 10 --warmup 3 --pt2-sources 0 --pt2-c4-sources 0 --no-cpu-baseline --conn-dets 0 --no-small-configs > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err
