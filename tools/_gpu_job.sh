set +e
timeout 900 python -m pytest tests/test_fullsize_properties_gpu.py -m gpu -x -q 2>&1 | tail -25 | cut -c1-300
