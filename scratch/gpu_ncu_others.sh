cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out /tmp/rep
timeout 600 ncu --set full --clock-control none -k regex:"k_levels_spec|k_prep_band|k_f32plane|k_fill_exact|k_exact|k_betab|k_filter_fast|k_bh" -c 16 -f -o /tmp/rep/others python scratch/prof_fast.py 20000 2 > gpurun_out/r02_other.log 2>&1; tail -1 gpurun_out/r02_other.log
ncu -i /tmp/rep/others.ncu-rep --page raw --csv > gpurun_out/r02_other_kernels_ncu_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"k_score_fast" -s 3 -c 3 -f -o /tmp/rep/union python scratch/prof_fast.py 20000 2 union > gpurun_out/r02_union.log 2>&1; tail -1 gpurun_out/r02_union.log
ncu -i /tmp/rep/union.ncu-rep --page raw --csv > gpurun_out/r02_k_score_fast_union_ncu_raw.csv 2>/dev/null
ls -la gpurun_out
