cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit code $?"; tail -30 gpurun_out/bench_n2.err; ls -la gpurun_out/bench_n2.json; head -c 600 gpurun_out/bench_n2.json
