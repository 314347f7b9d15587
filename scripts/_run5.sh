timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" -p no:cacheprovider 2>&1 | tail -5
python scripts/attn_probe.py
B200REC_ATTN_PIPE=0 python scripts/attn_probe.py
ncu --set full --clock-control none --import-source on -k "regex:attn_tc" --launch-skip 8 --launch-count 3 -o gpurun_out/attn_pipe python scripts/attn_probe.py > gpurun_out/attn_pipe.log 2>&1
ncu -i gpurun_out/attn_pipe.ncu-rep --page details > gpurun_out/attn_pipe_details.txt 2>/dev/null
ncu -i gpurun_out/attn_pipe.ncu-rep --page source --csv > gpurun_out/attn_pipe_source.csv 2>/dev/null
gzip -f gpurun_out/attn_pipe_source.csv; rm -f gpurun_out/attn_pipe.ncu-rep
