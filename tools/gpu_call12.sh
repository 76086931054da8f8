#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/c12_pytest.log 2>&1
tail -n 30 gpurun_out/c12_pytest.log
