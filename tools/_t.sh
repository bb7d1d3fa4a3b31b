python -m pytest tests -m gpu -x -q 2>&1 | tail -4
VARIANTS="default hs0" bash tools/ab_pq.sh 2>/dev/null | grep -v "^+"
