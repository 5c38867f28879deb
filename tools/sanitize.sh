#!/bin/bash
# compute-sanitizer passes over the tiny-model GPU tests of the kernels added last (memcheck, then racecheck on the kernels that
# exchange through shared memory / a ticket counter).  Round 1: 0 errors, 0 hazards.
cd "$(dirname "$0")/.."
export PYTHONPATH=.
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_qwen.py -m gpu -q -x \
  -k "case1 or batch_equals_single or split_attention or edges_vs_oracle" 2>&1 | tail -6
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_paraformer.py tests/test_gpu_sensevoice.py -m gpu -q -x \
  -k "batched_decoder or head128" 2>&1 | tail -6
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_qwen.py -m gpu -q -x -k "split_attention or tiled_prefill" 2>&1 | tail -6
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_paraformer.py -m gpu -q -x -k "batched_decoder and f32" 2>&1 | tail -6
