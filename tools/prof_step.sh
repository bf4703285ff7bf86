#!/bin/bash
# One bench step under ncu (run with gpurun): launch list, DRAM bytes of every GEMM launch, and --set full captures of
# the dominant kernels.  Outputs land in gpurun_out/; tools/summarize_launches.py / summarize_dram.py turn the csv
# files into the tables under profiles/.
TAG=${1:-r1}
mkdir -p gpurun_out
B="python bench.py --profile-step --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/p1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:gemm_tc --csv --log-file gpurun_out/gemm_dram_$TAG.csv $B > gpurun_out/p2.log 2>&1
# FFN w1 GEMM (bf16 + SiLU epilogue) is the 7th GEMM launch of a step; attention_tc: first launch
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc2 -s 6 -c 1 -o gpurun_out/prof_gemm_${TAG} -f $B > gpurun_out/p3.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc -c 1 -o gpurun_out/prof_attn_${TAG} -f $B > gpurun_out/p4.log 2>&1
ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/launches_$TAG.csv gpurun_out/gemm_dram_$TAG.csv
