mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_rowwise.py tests/test_gpu_blocks.py tests/test_gpu_lvdm.py -x -q 2>&1 | tail -15) > gpurun_out/s10_pytest.log
timeout 400 python tools/bench_rowwise.py > gpurun_out/s10_rowwise.jsonl 2> gpurun_out/s10_rowwise.err
VT_GN_TWOPASS=1 timeout 400 python tools/bench_rowwise.py 2>/dev/null | grep groupnorm > gpurun_out/s10_rowwise_gn_twopass.jsonl
timeout 400 python tools/bench_vc2_census.py > gpurun_out/s10_census.jsonl 2> gpurun_out/s10_census.err
cat gpurun_out/s10_pytest.log; tail -3 gpurun_out/s10_rowwise.err gpurun_out/s10_census.err
