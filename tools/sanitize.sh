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
# round 2: ragged batches (per-clip lengths read on the device), the FP8 instantiation of the streaming decode kernel, long-sequence attention
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_qwen_ragged.py -m gpu -q -x -k "generation_limit or (equals_single and bf16) or graph" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ragged.py tests/test_gpu_nar_ragged.py -m gpu -q -x -k "f32 or bf16" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fp8.py -m gpu -q -x -k "case0 or batch" 2>&1 | tail -6
# round 2, persistent Qwen decode-layer kernel (cooperative launch, grid barriers): memcheck 0 errors, racecheck 0 hazards
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_qwen_persist.py -m gpu -q -x -k "path and (2 or 4)" 2>&1 | tail -5
timeout 170 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_qwen_persist.py -m gpu -q -x -k "path and 2" 2>&1 | tail -6
# racecheck is not usable on decoder_stream_kernel: the only hazards it reports are the intentional volatile hand-off of the step
# counter between the worker warps and the MMA lanes (s_step / s_break_it, a monotonic flag that is polled), and the ~100x slow-down
# trips the kernel's own 3 s spin guard (`wait timed out`), bf16 and FP8 instantiation alike (round 2, gpurun_out/race_*.log).
