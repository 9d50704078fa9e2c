python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_step or packed_bound or fci" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest.log
python tools/ncu_dav.py 300 > gpurun_out/r02k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_dav' -s 60 -c 4 -o gpurun_out/r02k_dav python tools/ncu_dav.py 30 > gpurun_out/r02k_ncu.log 2>&1
