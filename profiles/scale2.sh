for w in wave_autograd_1e6 poisson_mat_4096; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --workload $w --no-cpu-baseline 2>gpurun_out/s16_scale2_$w.err | tail -1 > gpurun_out/s16_scale2_$w.json
python -c "
import json
d=json.loads(open('gpurun_out/s16_scale2_$w.json').read()); print('$w n=2 ms/step %.4f value %.4g e2e %.4g' % (d['ms_per_step'], d['value'], d['e2e']['value']))" || tail -5 gpurun_out/s16_scale2_$w.err
done
