#!/bin/bash
# usage: tools/gpu_n8_session.sh <tag> — bench at N=8 with the driver's launch line and default K/W
tag=$1; out=gpurun_out
t0=$SECONDS
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err
echo "n8 rc=$? $((SECONDS - t0)) s wall"
tail -3 $out/${tag}_bench_n8.err
python - $tag <<'PY'
import json,sys
d=json.loads(open('gpurun_out/%s_bench_n8.json' % sys.argv[1]).read().strip().splitlines()[-1])
e=d['e2e']; cg=e.get('by_channel_group') or {}
print('value',d['value'],'ms',d['ms_per_step'],'e2e',e['value'],e['mode'],'by_beam',e.get('by_beam'),'links',e['h2d_link_GBps_all_ranks'])
print('single',(cg.get('single_beam') or {}).get('value'),'all',(cg.get('all_beams') or {}).get('value'), cg.get('error'))
r=d.get('ring_e2e') or {}; print('ring',r.get('aggregate_GBps'),r.get('per_rank_GBps')); s=r.get('single_beam_one_stage_all_gpus') or {}; print('one stage', s.get('stage_GBps'), s.get('steady'), s.get('error'))
for t in (d.get('live_replay') or {}).get('trials',[]): print('live',{k:v for k,v in t.items() if k!='per_rank'})
print(d['clocks'], d['roofline']['frac'], d['beamset'])
PY
