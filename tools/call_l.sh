set -x
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "edge_shapes or big_path or deferred or non_pd" 2>&1 | tail -15
