#!/bin/bash
# ncu captures behind profiles/ncu_r01_summary.md (one GPU; a number printed under ncu is never a bench value)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
$NCU --set full --import-source on -k regex:decoder_ring -c 1 -f -o gpurun_out/prof_ring_r1 python tools/profile_ring.py 1 8 > gpurun_out/prof_ring.log 2>&1
$NCU --set full --import-source on -k regex:attention_tc -s 4 -c 1 -f -o gpurun_out/prof_attn_r1 python tools/profile_step.py > gpurun_out/prof_attn.log 2>&1
$NCU --set full --import-source on -k regex:gemm_tc -s 10 -c 4 -f -o gpurun_out/prof_gemm_r1 python tools/profile_step.py > gpurun_out/prof_gemm.log 2>&1
$NCU --set full -k regex:attention_tc -s 4 -c 1 -f -o gpurun_out/prof_attn_b4_r1 python tools/profile_step.py whisper-large-v3 4 > gpurun_out/prof_attn_b4.log 2>&1
tail -2 gpurun_out/prof_*.log
