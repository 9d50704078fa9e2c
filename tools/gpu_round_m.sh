python tools/ncu_projh4.py 4 > gpurun_out/r02m_plain.log 2>&1
