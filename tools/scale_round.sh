nproc > gpurun_out/r02s_nproc.txt; free -g | head -2 >> gpurun_out/r02s_nproc.txt
python bench.py --workload c5 --files 1000 --steps 2 --warmup 1 > gpurun_out/r02s_c5_n1.json 2> gpurun_out/r02s_c5_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload c5 --files 1000 --steps 2 --warmup 1 > gpurun_out/r02s_c5_n$n.json 2> gpurun_out/r02s_c5_n$n.err
done
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02s_c2_n$n.json 2> gpurun_out/r02s_c2_n$n.err
done
python bench.py --steps 10 > gpurun_out/r02s_c2_n1.json 2> gpurun_out/r02s_c2_n1.err
rm -rf /dev/shm/birda_b200_c5
