cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prep_band -s 2 -c 1 -o gpurun_out/prof_prep -f python scratch/e2e_probe2.py 2 > gpurun_out/b_ncu3.log 2>&1
