python -m pytest tests/test_gpu_predict.py -m gpu -q 2>&1 | tail -25
