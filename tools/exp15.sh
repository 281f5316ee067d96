for sw in 1 0; do
for shape in 128,64,64,128 10,256,256,256 16,512,512,64 1,512,512,100; do
    echo "== sweep=$sw shape=$shape"
    BDOF_SWEEP=$sw python bench.py --steps 3 --warmup 3 --no-cpu --shape $shape 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f'%d['value'], {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    else: print(l.strip()[:300])
"
done
done
