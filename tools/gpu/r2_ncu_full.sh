#!/bin/bash
# round-2: ncu --set full captures of the two conv kernels on the final build (one launch each, one cfg2 B=64 forward)
cd "$(dirname "$0")/../.."
timeout 170 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_conv_umma -s 1 -c 1 -f -o gpurun_out/r2_ncu_conv_umma_final python tools/profile_forward.py cfg2 64 > gpurun_out/r2_ncu_a.log 2>&1
timeout 170 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_conv1x1 -s 2 -c 1 -f -o gpurun_out/r2_ncu_conv1x1_final python tools/profile_forward.py cfg2 64 > gpurun_out/r2_ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
python tools/ncu_brief.py gpurun_out/r2_ncu_conv_umma_final.ncu-rep gpurun_out/r2_ncu_conv1x1_final.ncu-rep > gpurun_out/r2_ncu_final_brief.txt 2>&1
cat gpurun_out/r2_ncu_final_brief.txt | head -32
