python -m pytest tests/test_gpu_parity.py -m gpu -q -k "edge_shapes" 2>&1 | tail -15
