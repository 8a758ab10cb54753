python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 4 --steps 5 > gpurun_out/s9_c4_n2.json 2> gpurun_out/s9_c4_n2.err; tail -c 1800 gpurun_out/s9_c4_n2.json; tail -5 gpurun_out/s9_c4_n2.err
