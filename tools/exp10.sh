for sg in 0 300 700 1200; do
for shape in 1,4096,4096,48 1,2048,2048,128; do
    echo "== stagger_ns=$sg shape=$shape"
    BDOF_STAGGER_NS=$sg python bench.py --steps 3 --warmup 3 --no-cpu --shape $shape 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f'%d['value'], {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    else: print(l.strip()[:300])
"
done
done
