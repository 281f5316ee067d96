for tune in 0 1 3 7; do
  for shape in 1,4096,4096,48 1,2048,2048,128; do
    echo "== tune=$tune shape=$shape"
    BDOF_TUNE=$tune python bench.py --steps 3 --warmup 3 --no-cpu --shape $shape 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f'%d['value'], {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    else: print(l.strip()[:200])
"
  done
done
