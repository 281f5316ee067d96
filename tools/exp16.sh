for shape in 1,1024,1024,128 4,1024,1024,64 1,512,512,100 10,256,256,256; do
    echo "== shape=$shape"
    python bench.py --steps 3 --warmup 3 --no-cpu --shape $shape 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f'%d['value'], {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    else: print(l.strip()[:300])
"
done
