python tools/ncu_dav.py 30 > gpurun_out/r02k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_dav' -s 60 -c 6 -o gpurun_out/r02k_dav python tools/ncu_dav.py 30 > gpurun_out/r02k_ncu.log 2>&1
