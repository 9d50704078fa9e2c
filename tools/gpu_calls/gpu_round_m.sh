python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed or guard" > gpurun_out/r02m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02m_pytest.log
python tools/ncu_projh4.py 4 > gpurun_out/r02m_plain.log 2>&1
