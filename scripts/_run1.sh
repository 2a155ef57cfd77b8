timeout 900 python -m pytest tests/test_gpu_observables.py -x -q -s 2>&1 | tail -20
