python tools/exp_spmv_block.py >> gpurun_out/r02o_spmv_block.log 2>&1
