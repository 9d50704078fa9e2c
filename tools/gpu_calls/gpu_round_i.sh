# compute-sanitizer memcheck over the shipped kernels on a small case (one tool per call)
python tools/ncu_targets.py --small > gpurun_out/r02i_plain.log 2>&1 &&
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 python tools/ncu_targets.py --small > gpurun_out/r02i_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02i_memcheck.log
