R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8"
$R --config 4 --steps 5 > gpurun_out/s10_c4_n8.json 2> gpurun_out/s10_c4_n8.err; tail -c 600 gpurun_out/s10_c4_n8.json
$R --config 5 --steps 5 > gpurun_out/s10_c5_n8.json 2> gpurun_out/s10_c5_n8.err; tail -c 300 gpurun_out/s10_c5_n8.json
$R --config 2 --steps 20 > gpurun_out/s10_c2_n8.json 2> gpurun_out/s10_c2_n8.err; tail -c 300 gpurun_out/s10_c2_n8.json
python bench.py --impl reference --config 5 > gpurun_out/s10_ref_c5.json 2>/dev/null
