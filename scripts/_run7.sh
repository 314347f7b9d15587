export REPS=4
ncu --set full --clock-control none --import-source on -k "regex:attn_tc" --launch-skip 4 --launch-count 1 -o gpurun_out/attn_f python scripts/attn_probe.py > gpurun_out/attn_f.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:attn_tc" --launch-skip 9 --launch-count 2 -o gpurun_out/attn_b python scripts/attn_probe.py > gpurun_out/attn_b.log 2>&1
for f in attn_f attn_b; do
ncu -i gpurun_out/$f.ncu-rep --page details > gpurun_out/${f}_details.txt 2>/dev/null
ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/${f}_source.csv 2>/dev/null
gzip -f gpurun_out/${f}_source.csv; rm -f gpurun_out/$f.ncu-rep
done
