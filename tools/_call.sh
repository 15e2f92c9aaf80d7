python -m pytest tests/test_gpu_points.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4
for z in 1 2; do python bench.py --grid 128 --zoom $z --steps 30 | tail -1; done
python tools/time_lists.py 256 2>&1 | grep "overlap=True"
