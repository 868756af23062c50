for v in default nodiscard default nodiscard; do
  if [ $v = default ]; then unset MMF_LIB_PATH; else export MMF_LIB_PATH=$PWD/multimodalfusion_b200/libmmf_b200_nodiscard.so; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 48 --warmup 8 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['ms_per_step']*1e3,2), 'us/step', round(d['value']/1e6,1), 'M')
"
done
