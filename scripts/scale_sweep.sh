set -x
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 8 --warmup 3 --no-e2e --no-alt --no-cpu-baseline --slices $2 2>> gpurun_out/r2b_scale_sweep.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'slices': '$2', 'env': '$3', 'value': d['value'], 'ms_per_step': d['ms_per_step'], 'kernel_ms': d['roofline']['kernel_ms'], 'ok': d['multi_gpu']['gathered_equals_single_gpu_rows']}))" >> gpurun_out/r2b_scale_sweep.jsonl; }
rm -f gpurun_out/r2b_scale_sweep.jsonl
run 29511 0.6,0.4 -
run 29512 0.5,0.3,0.2 -
run 29513 0.4,0.3,0.2,0.1 -
run 29514 0.45,0.3,0.15,0.1 -
NCCL_MAX_CTAS=4 run 29515 0.5,0.3,0.2 NCCL_MAX_CTAS=4
cat gpurun_out/r2b_scale_sweep.jsonl
