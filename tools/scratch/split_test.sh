# GPU validation after a kernel change: the GPU test-suite, then the shape sweep of the split forms, the sanitizer cases, the size sweep
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/split_shapes.py 2>&1 | tee gpurun_out/r2_split_shapes.log | tail -20
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_cases.py 2>&1 | tail -4
timeout 200 python tools/size_sweep.py 2>&1 | head -12
