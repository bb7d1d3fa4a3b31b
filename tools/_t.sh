python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/perf_scan.py 1,2,3,4 1000000 2>&1 | grep -v Warn
