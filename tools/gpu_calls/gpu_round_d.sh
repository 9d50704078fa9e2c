python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest.log
python tools/exp_small.py > gpurun_out/r02d_exp_small.jsonl 2> gpurun_out/r02d_exp_small.err
