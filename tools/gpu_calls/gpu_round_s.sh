python tools/exp_small.py rounds > gpurun_out/r02s_rounds.jsonl 2> gpurun_out/r02s_rounds.err
python bench.py --steps 10 --warmup 3 --pt2-sources 0 --pt2-c4-sources 0 --skqd-nf 0 --no-cpu-baseline --conn-dets 0 > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
