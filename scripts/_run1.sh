timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_holstein.py -m gpu -x -q --durations=5 -k "not test_B_p and not test_E_p" 2>&1 | tail -25
