mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_pipeline_gpu.py -k "npz" -m gpu -q -s -x > gpurun_out/r02m_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r02m_tests.log
