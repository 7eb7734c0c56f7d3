set -x
PYTHONPATH=. timeout 600 python profiles/scripts/time_discover.py 20000 10 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_scan_gpu.py -x -q -m gpu -k "tensor_path_zero or tensor_path_short or config1" 2>&1 | tail -15
