python tools/ncu_targets2.py > gpurun_out/r02u_plain.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:k_projh4|k_dav|k_taylor|k_lists|k_peer_gather|k_peer_step|k_conn' -c 30 -o gpurun_out/r02u_final python tools/ncu_targets2.py > gpurun_out/r02u_ncu.log 2>&1
ls -la gpurun_out > gpurun_out/r02u_ls.log
