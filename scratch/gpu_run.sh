cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "n129 or n500_b60 or n515 or n400" 2>&1 | tail -12
echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scratch/dbg.py spec 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_apa.py tests/test_gpu_golden.py -x -q -m gpu -k "apa_rejections or bhfdr_matches_reference and synth_p2w5 or prep and synth_p2w5" 2>&1 | tail -6
