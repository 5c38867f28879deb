#!/bin/bash
# Runs the GPU parity suites in separate processes (a trapped kernel kills its CUDA context,
# so one bad suite must not take the others down) with hard timeouts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
for t in ${SUITES:-test_gpu_fullsize test_gpu_edges test_gpu_qwen test_gpu_paraformer test_gpu_sensevoice test_gpu_protocol test_gpu_attention test_gpu_ring test_gpu_gemm test_gpu_whisper_f32 test_gpu_whisper_bf16}; do
  echo "=== $t"
  timeout 900 python -m pytest tests/$t.py -m gpu -q -s --timeout 600 2>&1 | tail -40 | tee gpurun_out/$t.log
done
