timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/bench_ground.py 512 2 2>&1 | tail -3 | tee gpurun_out/bench_ground.json
