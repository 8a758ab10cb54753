R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4"
$R --config 4 --steps 5 > gpurun_out/s10_c4_n4.json 2> gpurun_out/s10_c4_n4.err; tail -c 600 gpurun_out/s10_c4_n4.json
