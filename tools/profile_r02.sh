#!/bin/bash
# ncu captures behind profiles/ncu_r02_stream_summary.md (one GPU; a number printed under ncu is never a bench value)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r02.csv python tools/profile_step.py > gpurun_out/prof_step_r02.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:decoder_stream -c 1 -f -o gpurun_out/prof_stream_b1_r02 python tools/profile_ring.py 1 8 > gpurun_out/prof_stream_b1.log 2>&1
timeout 900 $NCU --set full -k regex:decoder_stream -c 1 -f -o gpurun_out/prof_stream_b4_r02 python tools/profile_ring.py 4 8 > gpurun_out/prof_stream_b4.log 2>&1
timeout 900 $NCU --set full -k regex:decoder_stream -c 1 -f -o gpurun_out/prof_stream_b1_fp8_r02 python tools/profile_ring.py 1 8 1 > gpurun_out/prof_stream_b1_fp8.log 2>&1
timeout 900 $NCU --set full -k regex:attention_tc -c 1 -f -o gpurun_out/prof_attn_b4_r02 python tools/profile_step.py whisper-large-v3 4 > gpurun_out/prof_attn_b4.log 2>&1
timeout 900 $NCU --set full -k regex:gemm_tc -c 3 -f -o gpurun_out/prof_gemm_b4_r02 python tools/profile_step.py whisper-large-v3 4 > gpurun_out/prof_gemm_b4.log 2>&1
tail -2 gpurun_out/prof_step_r02.log gpurun_out/prof_stream_b1.log gpurun_out/prof_stream_b4.log
ls -la gpurun_out/*.ncu-rep
