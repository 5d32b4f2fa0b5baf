# usage: bash tools/run_one.sh "<pytest -k expression>"
timeout 900 python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | tail -25
