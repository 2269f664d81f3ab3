timeout 1500 python -m pytest tests/test_ukf_gpu.py tests/test_closed_loop_gpu.py tests/test_user_model_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/dev_time_ukf.py 2>&1 | grep "upd/s"
