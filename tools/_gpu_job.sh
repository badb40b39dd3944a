set +e
timeout 200 python -m pytest tests/test_dbn.py -m gpu -q 2>&1 | tail -30
