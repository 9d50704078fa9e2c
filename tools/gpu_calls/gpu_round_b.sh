# round-2 GPU call B: tests, then ncu over the shipped kernels (launch list + one --set full capture)
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
python tools/ncu_targets.py > gpurun_out/r02b_targets_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches.csv python tools/ncu_targets.py > gpurun_out/r02b_ncu_launches.log 2>&1
python tools/ncu_targets.py > gpurun_out/r02b_targets_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_pt2_accumulate2|k_pt2_score|k_projh3|k_spmv_sell_f32|k_peer_step|k_sell_fill|k_sell_pack' -c 14 -o gpurun_out/r02b_shipped python tools/ncu_targets.py > gpurun_out/r02b_ncu_full.log 2>&1
