python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02r_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02r_smoke.log
