#!/bin/bash
# usage: tools/gpu_scale_session.sh <tag> — on an 8-GPU box: bench at N=8, N=4 (spread), N=4 (identity, e2e only)
tag=$1
out=gpurun_out
{ nproc; free -g | head -2; nvidia-smi topo -m; nvidia-smi --query-gpu=index,pci.bus_id --format=csv; } > $out/${tag}_box.txt 2>&1
run() { # n port extra...
  n=$1; port=$2; name=$3; shift 3
  t0=$SECONDS
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 "$@" > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err
  echo "$name rc=$? $((SECONDS - t0)) s wall"
}
run 8 29521 n8
run 4 29522 n4
run 2 29523 n2
python bench.py --impl reference > $out/${tag}_ref.json 2> $out/${tag}_ref.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python -  <<'PY'
import json,glob,sys
tag=sys.argv[1] if len(sys.argv)>1 else ''
for f in sorted(glob.glob('gpurun_out/%s_n*.json' % tag)):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'unreadable',e); continue
    e=d['e2e']; cg=e.get('by_channel_group') or {}
    print(f, 'value',d['value'],'e2e',e['value'],e['mode'],'by_beam',e.get('by_beam'),'links',e['h2d_link_GBps_all_ranks'],'sum',e['h2d_links_sum_GBps'])
    print('   cg chunks',cg.get('chunks_per_gpu'),'single',(cg.get('single_beam') or {}).get('value'),'all',(cg.get('all_beams') or {}).get('value'), cg.get('error'))
    r=d.get('ring_e2e') or {}; print('   ring',r.get('aggregate_GBps'),r.get('per_rank_GBps'))
    l=d.get('live_replay') or {}
    for t in l.get('trials',[]): print('   live',{k:v for k,v in t.items() if k!='per_rank'})
    print('   placement',d['placement'])
PY
