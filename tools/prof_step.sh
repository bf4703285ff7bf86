#!/bin/bash
# One bench step under ncu (run with gpurun): launch list, DRAM bytes of every GEMM launch, and --set full captures of
# the front of the path (fbank, conv1, conv2, embed), one whole encoder layer (16 launches) and the search kernels.
# Outputs land in gpurun_out/; tools/summarize_launches.py / summarize_dram.py / ncu_key_metrics.py turn them into the
# tables under profiles/.
TAG=${1:-r1}
mkdir -p gpurun_out
B="python bench.py --profile-step --no-cpu-baseline"
N="ncu --clock-control none --profile-from-start off"
$N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/p1.log 2>&1
$N --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_tc --csv --log-file gpurun_out/gemm_dram_$TAG.csv $B > gpurun_out/p2.log 2>&1
# launch order of a step: 0 fbank, 1 linear_pos GEMM, 2 conv1, 3 conv2 GEMM, 4 embed GEMM, 5.. encoder layers;
# launches 23..39 = [final LN of layer 0 +] the 16 kernels of encoder layer 1
$N --set full --import-source on -c 5 -o gpurun_out/prof_front_$TAG -f $B > gpurun_out/p3.log 2>&1
$N --set full --import-source on -s 23 -c 17 -o gpurun_out/prof_layer_$TAG -f $B > gpurun_out/p4.log 2>&1
$N --set full --import-source on -k "regex:logsoftmax_topk|ctc_prefix_beam|logsoftmax_gather" -c 3 -o gpurun_out/prof_search_$TAG -f $B > gpurun_out/p5.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep; wc -l gpurun_out/launches_$TAG.csv gpurun_out/gemm_dram_$TAG.csv
