python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for shape in 1,4096,4096,48 1,2048,2048,128 1,1024,1024,128 4,2048,2048,32; do
    echo "== shape=$shape"
    python bench.py --steps 3 --warmup 3 --no-cpu --shape $shape 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f'%d['value'], {k:round(v['avg_ms']*1e3,1) for k,v in d['kernels'].items()})
    else: print(l.strip()[:300])
"
done
BDOF_LIB=libbdof_alt9.so python tools/phase_sweep.py 4096 | grep -E "^==|tile 1"
BDOF_LIB=libbdof_alt9.so python tools/phase_sweep.py 2048 | grep -E "^==|tile 1"
