python -m pytest tests/test_gpu_large.py -m gpu -x -q -k "d2 or splits" 2>&1 | tail -3
python tools/perf_scan.py 2 1000000 2>&1 | grep -v Warn
