# C5 (1000 files in /dev/shm) at N = 8, 4, 2, 1 ranks on one box; files are written once
python bench.py --workload c5 --files 1000 --steps 2 --warmup 1 > gpurun_out/r02t_c5_n1.json 2> gpurun_out/r02t_c5_n1.err
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --workload c5 --files 1000 --steps 2 --warmup 1 > gpurun_out/r02t_c5_n$n.json 2> gpurun_out/r02t_c5_n$n.err
done
rm -rf /dev/shm/birda_b200_c5
