python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_v7.log 2>&1; tail -3 gpurun_out/pytest_gpu_v7.log
python scripts/middlebury_study.py 30000 all > gpurun_out/middlebury_r01.log 2>&1; cat gpurun_out/middlebury_r01.log | cut -c1-400
