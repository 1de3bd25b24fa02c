#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_train_golden_gpu.py -m gpu -q -x -s 2>&1 | tail -25 > gpurun_out/r2z_pytest.txt
tail -25 gpurun_out/r2z_pytest.txt
timeout 600 python -c "
import bench, torch, json
print(json.dumps(bench.cv_stage(torch.device('cuda', 0), 0, 1)))" > gpurun_out/r2z_cv.json 2> gpurun_out/r2z_cv.err
cat gpurun_out/r2z_cv.json; tail -5 gpurun_out/r2z_cv.err
