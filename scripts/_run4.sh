timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" -p no:cacheprovider 2>&1 | tail -5
timeout 300 python bench.py --config C --profile --steps 5 --warmup 3 --no-cpu --no-eval --sustained-seconds 0 > gpurun_out/bench_profile_C2.json 2> gpurun_out/bench_profile_C2.txt; echo "profile C rc=$?"
grep -E "attn_tc|Self CUDA time" gpurun_out/bench_profile_C2.txt | cut -c1-60,100-200
B200REC_ATTN_PIPE=0 timeout 300 python bench.py --config C --profile --steps 5 --warmup 3 --no-cpu --no-eval --sustained-seconds 0 > gpurun_out/bench_profile_C2_old.json 2> gpurun_out/bench_profile_C2_old.txt
grep -E "attn_tc|Self CUDA time" gpurun_out/bench_profile_C2_old.txt | cut -c1-60,100-200
