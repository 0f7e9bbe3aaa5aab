#!/bin/bash
N=$1; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo bench rc=$?
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n$N.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['lm_iter']['s_per_iter'], d['lm_iter']['stage_ms_rank0'], d['lm_iter']['step_check'], d['multi_gpu_parity']['ok'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
"
timeout 900 $TR --master-port 29513 tools/sweep.py "$@" > gpurun_out/r2_sweep_n$N.jsonl 2> gpurun_out/r2_sweep_n$N.err; echo sweep rc=$?
cut -c1-330 gpurun_out/r2_sweep_n$N.jsonl
