python -m pytest tests/test_gpu_shared.py -x -q 2>&1 | tail -3
python tools/time_lists.py 256 2>&1 | grep "overlap=True"
python tools/time_lists.py 128 2>&1 | grep "overlap=True"
python tools/time_lists.py 64 2>&1 | grep "overlap=True"
