#!/bin/bash
# usage: tools/gpu_probe_check.sh <tag> <ngpu> — short bench (e2e legs only) + one group stage over the GPUs
tag=$1; n=$2; out=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 8 --e2e-steps 6 --no-ring --no-live --no-cpu --beamset 0 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
echo "bench rc=$?"; tail -2 $out/${tag}_bench_n$n.err
python - $tag $n <<'PY'
import json,sys
d=json.loads(open('gpurun_out/%s_bench_n%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
e=d['e2e']; cg=e.get('by_channel_group') or {}
print('e2e',e['value'],e['mode'],'by_beam',e.get('by_beam'))
print('links together',e['h2d_link_GBps_all_ranks'],'sum',e['h2d_links_sum_GBps'],'equal-bytes pass',e['h2d_link_GBps_equal_bytes_pass'], 'frac', e['frac_of_h2d_links_sum'])
for k in ('single_beam','all_beams'):
    x=cg.get(k) or {}; print(k, x.get('value'), x.get('units_per_gpu'), len(x.get('refinement') or []), 'refinement passes', cg.get('error'))
PY
g=$(python -c "print(','.join(str(i*(8//$n) if $n<8 else i) for i in range($n)))")
[ $(nvidia-smi -L | wc -l) -lt 8 ] && g=$(python -c "print(','.join(str(i) for i in range($n)))")
OMP_NUM_THREADS=$(nproc) timeout 200 python tools/run_ring_e2e.py --gpu $g --nblocks 64 --nbufs 3 > $out/${tag}_group_$n.json 2> $out/${tag}_group_$n.err; echo "group rc=$?"; cat $out/${tag}_group_$n.json; tail -2 $out/${tag}_group_$n.err
