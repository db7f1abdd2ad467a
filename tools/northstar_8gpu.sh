# BASELINE configs[2]: the 4096-view Room sweep on 8 GPUs (strong-scaled: 512 views per GPU), reference-parity scorer and NGP mode,
# next to the host-CPU reference on the same box.  Lines go to gpurun_out/northstar_*.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 8 --workload room --steps 2 --warmup 3 > gpurun_out/northstar_room_n8.json 2> gpurun_out/northstar_room_n8.err
timeout 300 $TR bench.py --gpus 8 --mode ngp --steps 5 --warmup 3 > gpurun_out/northstar_ngp_n8.json 2> gpurun_out/northstar_ngp_n8.err
timeout 300 python bench.py --impl reference --workload room --steps 1 --warmup 1 > gpurun_out/northstar_room_ref.json 2> gpurun_out/northstar_room_ref.err
grep -h '^{' gpurun_out/northstar_*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d.get('impl', 'evpp'), d['n_gpus'], d['value'], d['e2e']['value'], d['config']['workload'][:60], d.get('cpu_baseline', {}).get('cores'))
"
