python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/perf_scan.py 1,2,3,4 1000000 2>&1 | grep -v Warn
python tools/perf_scan.py 1,2,3,4 100000 2>&1 | grep -v Warn
