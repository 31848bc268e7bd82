cd $GRAFT_REPO_ROOT
for rep in 1 2 3 4 5 6 7 8 9 10 11 12; do
  for v in noflags; do
    E="OSB_NO_ROW_FLAGS=1"
    env $E timeout -s USR1 -k 10 50 python bench.py --steps 300 --warmup 5 --no-oscr --no-sub --no-e2e > gpurun_out/r03c_${v}_$rep.json 2> gpurun_out/r03c_${v}_$rep.err
    rc=$?; echo "$v $rep rc=$rc"; if [ $rc -ne 0 ]; then tail -40 gpurun_out/r03c_${v}_$rep.err; fi
  done
done
