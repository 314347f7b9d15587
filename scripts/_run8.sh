timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" -p no:cacheprovider 2>&1 | tail -3
timeout 120 python scripts/attn_probe.py
timeout 300 python bench.py --config C --steps 10 --warmup 3 --no-cpu --no-eval --sustained-seconds 3 > gpurun_out/bench_C.json 2> gpurun_out/bench_C.err; echo "bench C rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_C.json").read().strip().splitlines()[-1])
print("C:", d["value"], d["ms_per_step"], d.get("sustained",{}).get("value"), d["roofline"]["achieved"])
PY
