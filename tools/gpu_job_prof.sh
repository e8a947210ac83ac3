#!/bin/bash
# ncu evidence on the final code: per-launch list of one C2 step + full captures of the dominant kernels
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_C2.csv python tools/profile_step.py --steps 2 > $O/launches_C2.out 2>&1
python tools/summarize_launches.py $O/launches_C2.csv 2 > $O/launches_C2.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|convt4_kernel" -c 12 -o $O/conv_full python tools/profile_conv_gemm.py > $O/conv_full.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_cl" -c 4 -o $O/lstm_full python tools/bench_lstm_scan_once.py > $O/lstm_full.out 2>&1
timeout 300 python tools/profile_step.py --steps 3 --calls > $O/calls_C2.txt 2>&1
timeout 400 python tools/profile_step.py --steps 3 --calls --backbone vgg_64 --channels 3 --batch 128 > $O/calls_C3.txt 2>&1
R=512 timeout 200 python tools/bench_lstm_scan.py 2>&1 | grep -v "tf32=0" > $O/scan_R512.txt
ls -la $O
