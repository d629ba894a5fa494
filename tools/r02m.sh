python -m pytest tests/test_cascade.py -m gpu -x -q 2>&1 | tail -15
python tools/cascade_probe.py 400000
