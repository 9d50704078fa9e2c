for s in 0 8 19 48 96 192 384; do
  echo "FGK_PT2_SPLIT=$s" >> gpurun_out/r02n_split.log
  FGK_PT2_SPLIT=$s python tools/exp_pt2_cap.py 114000000 >> gpurun_out/r02n_split.log 2>&1
done
