timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm" -p no:cacheprovider 2>&1 | tail -3
SKIP_TESTS=1 bash scripts/gpu_round_check.sh 2>&1 | tail -18
timeout 300 python bench.py --config C --profile --steps 3 --warmup 3 --no-cpu --no-eval --sustained-seconds 0 > gpurun_out/bench_profile_C.json 2> gpurun_out/bench_profile_C.txt; echo "profile C rc=$?"
bash scripts/ncu_kernels.sh r02c_a "gate_ln|layernorm_bwd|attn_seq" 108 10
bash scripts/ncu_kernels.sh r02c_b "nce_pos|nce_combine|l2norm|colsum_partial" 106 14
bash scripts/ncu_kernels.sh r02c_c "nce_pos|nce_combine|l2norm|colsum_partial" 159 8
