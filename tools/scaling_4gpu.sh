# Fills the 1/2/4/8 table: round 1 Object workload (weak) at N = 4, the Room sweep (the default since round 2, strong) at N = 2 and 4.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 4 --master-port 29551 bench.py --gpus 4 --workload object --steps 3 --warmup 3 > gpurun_out/scale_object_n4.json 2> gpurun_out/scale_object_n4.err
timeout 300 $TR --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --workload room --steps 2 --warmup 3 > gpurun_out/scale_room_n4.json 2> gpurun_out/scale_room_n4.err
timeout 300 $TR --nproc-per-node 2 --master-port 29553 bench.py --gpus 2 --workload room --steps 2 --warmup 3 > gpurun_out/scale_room_n2.json 2> gpurun_out/scale_room_n2.err
grep -h '^{' gpurun_out/scale_*_n*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['workload'][:40])
"
