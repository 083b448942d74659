#!/bin/bash
# Run on the GPU box (under gpurun): ncu captures for the kernels + launch list of bench.py.
# usage: tools/gpu_profile_round.sh TAG
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import birda_b200 as b; c=b.Context(0); print(b.FrontEndPlan(c,44100,2,b.FMT_S16,48000,144000,72000).describe())" > gpurun_out/${TAG}_k2_describe.txt
ncu --set full --clock-control none --import-source on -k regex:resample_ -s 2 -c 1 -f -o gpurun_out/${TAG}_k2_c2_1h python tools/prof_run.py 60 2 c2 > gpurun_out/${TAG}_ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resample_ -s 2 -c 1 -f -o gpurun_out/${TAG}_k2_c3_1h python tools/prof_run.py 60 2 c3 > gpurun_out/${TAG}_ncu_k2c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pack_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_k1_c4 python tools/prof_run.py 20 2 c4 > gpurun_out/${TAG}_ncu_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_k3_c2 python tools/prof_run.py 60 2 c2 > gpurun_out/${TAG}_ncu_k3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_power -s 2 -c 1 -f -o gpurun_out/${TAG}_k5a_low python tools/prof_mel.py 594 2 low > gpurun_out/${TAG}_ncu_k5a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mel_gemm -s 2 -c 1 -f -o gpurun_out/${TAG}_k5b_low python tools/prof_mel.py 594 2 low > gpurun_out/${TAG}_ncu_k5b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mel_gemm -s 2 -c 1 -f -o gpurun_out/${TAG}_k5b_full python tools/prof_mel.py 594 2 full > gpurun_out/${TAG}_ncu_k5bf.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"resample|pack_kernel|post_kernel|stft_power|mel_gemm" -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
for f in k2 k1 k3 k5a k5b; do tail -1 gpurun_out/${TAG}_ncu_$f.log; done
