timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" -p no:cacheprovider 2>&1 | tail -5
timeout 120 python scripts/attn_probe.py
