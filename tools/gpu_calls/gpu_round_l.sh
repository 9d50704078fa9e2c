python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02l_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r02l_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02l_ref.json 2> gpurun_out/r02l_ref.err; echo "ref exit $?" >> gpurun_out/r02l_ref.err
python tools/ncu_projh4.py > gpurun_out/r02l_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:k_projh4|k_lists' -c 6 -o gpurun_out/r02l_projh4 python tools/ncu_projh4.py > gpurun_out/r02l_ncu.log 2>&1
