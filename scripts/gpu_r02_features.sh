mkdir -p gpurun_out/r02feat
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_adaptor_under_reference.py tests/test_gpu_learner.py -m gpu -q -x -k "features or incremental or column_tile or cuda_balloon_arena or eval or device or train or quantile" > gpurun_out/r02feat/pytest_feat.log 2>&1; tail -3 gpurun_out/r02feat/pytest_feat.log
for rep in 1 2; do timeout 300 python scripts/feature_timing.py --num-envs 65536 2>&1 | tail -1 | tee -a gpurun_out/r02feat/feature_timing.jsonl; done
