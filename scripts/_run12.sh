B200REC_PROFILE_EVAL=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --sustained-seconds 0 --instrument-steps 0 2> gpurun_out/eval_profile.txt > gpurun_out/eval_profile.json
grep -v "^-" gpurun_out/eval_profile.txt | cut -c1-62,112-200 | head -24
