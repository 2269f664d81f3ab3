WSDBG=0 CQS=5 timeout 300 python tools/dev_ws.py c1 1 2>&1 | tail -2
timeout 300 python tools/dev_timeline.py 2>&1 | head -20
timeout 900 python -m pytest tests/test_mppi_gpu.py -x -q -m gpu 2>&1 | tail -3
