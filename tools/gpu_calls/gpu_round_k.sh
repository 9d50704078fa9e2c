python bench.py --steps 10 --warmup 3 --pt2-sources 0 --pt2-c4-sources 0 --skqd-nf 0 --no-cpu-baseline --conn-dets 0 --krylov-phases > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
