mkdir -p gpurun_out
B200REC_PROFILE_STEP=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 5 --no-cpu --no-eval --sustained-seconds 0 --instrument-steps 0 > gpurun_out/bench_2gpu_prof.json 2> gpurun_out/bench_2gpu_prof.err; echo "rc=$?"
ls -la gpurun_out/trace_rank0.json; gzip -f gpurun_out/trace_rank0.json
tail -c 600 gpurun_out/bench_2gpu_prof.json
