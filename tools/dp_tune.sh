# N-GPU data-parallel tuning sweep: NCCL CTA cap x SMs reserved for NCCL during the overlapped backward
N=${1:-2}
port=29560
for cfg in ${DP_CFGS:-"0:0 16:16 8:8 16:0 32:32"}; do cfg=${cfg/:/ }; set -- $cfg; port=$((port+1)); timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --nccl-max-ctas $1 --comm-sm-reserve $2 2> gpurun_out/dp_${N}_$1_$2.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ctas/reserve', '$1', '$2', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['dp_mode'])"; done
