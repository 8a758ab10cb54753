O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
$R bench.py --gpus 8 --config 2 --steps 20 > $O/g8_c2_n8.json 2> $O/g8_c2_n8.err; tail -c 300 $O/g8_c2_n8.json
$R bench.py --gpus 8 --config 4 --steps 5 > $O/g8_c4_n8.json 2> $O/g8_c4_n8.err; tail -c 600 $O/g8_c4_n8.json
$R bench.py --gpus 8 --config 5 --steps 5 > $O/g8_c5_n8.json 2> $O/g8_c5_n8.err; tail -c 300 $O/g8_c5_n8.json
$R scripts/d2h_concurrency.py > $O/g8_d2h.json 2> $O/g8_d2h.err; tail -c 1500 $O/g8_d2h.json
R4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513"
CUDA_VISIBLE_DEVICES=0,1,2,3 $R4 bench.py --gpus 4 --config 4 --steps 5 > $O/g8_c4_n4.json 2> $O/g8_c4_n4.err; tail -c 300 $O/g8_c4_n4.json
R2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
CUDA_VISIBLE_DEVICES=0,1 $R2 bench.py --gpus 2 --config 4 --steps 5 > $O/g8_c4_n2.json 2> $O/g8_c4_n2.err; tail -c 300 $O/g8_c4_n2.json
