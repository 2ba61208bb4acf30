set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/profile_train.py 2 fp32 C4 --table > gpurun_out/train_table_c4.txt 2>&1; tail -50 gpurun_out/train_table_c4.txt
python tools/profile_train.py 2 fp32 C2_TRAIN --table > gpurun_out/train_table_c2.txt 2>&1; tail -45 gpurun_out/train_table_c2.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1k_synth_launches.csv python tools/profile_step.py 1 fp32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 4 -c 1 -o gpurun_out/r1k_attention_tc python tools/profile_step.py 1 fp32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 3 -o gpurun_out/r1k_gemm_tc python tools/profile_step.py 1 fp32 > /dev/null 2>&1
ls -la gpurun_out/
