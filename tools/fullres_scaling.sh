# BASELINE configs[4] (full-resolution render sweep, ray-sharded): N = 1 and N = all visible GPUs
N=$(nvidia-smi -L | wc -l)
timeout 300 python bench.py --mode fullres --steps 3 --warmup 3 > gpurun_out/fullres_n1.json 2> gpurun_out/fullres_n1.err
if [ "$N" -gt 1 ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --mode fullres --steps 3 --warmup 3 > gpurun_out/fullres_n$N.json 2> gpurun_out/fullres_n$N.err
fi
for f in gpurun_out/fullres_n*.err; do tail -n 3 $f; done
grep -h '^{' gpurun_out/fullres_n*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], 'views/s', d['value'], 'e2e', d['e2e']['value'], 'rays/s', d['rays_per_s'], 'MLP TF/s', d['roofline']['achieved'], 'share', d['roofline']['mlp_share_of_step'], 'launches', d['gpu_launches'], d['check'])
"
