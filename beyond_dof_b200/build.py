"""Build libbdof.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m beyond_dof_b200.build [--force]

One translation unit per FFT length (csrc/line_inst.cu, -DBDOF_N=...) plus csrc/bdof.cu, compiled
in parallel and linked into beyond_dof_b200/libbdof.so.  Objects are cached in csrc/_obj/ and
rebuilt when any source is newer.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_obj')
ALT = int(os.environ.get('BDOF_ALT', '0'))          # experimental FFT factorizations (csrc/line_inst.cu)
LIB = os.path.join(HERE, 'libbdof.so' if ALT == 0 else 'libbdof_alt%d.so' % ALT)
if ALT:
    OBJ = os.path.join(CSRC, '_obj_alt%d' % ALT)
SIZES = [64, 128, 256, 512, 1024, 2048, 4096, 8192]
NVCC_FLAGS = ['-O3', '-std=c++17', '--expt-relaxed-constexpr', '-gencode', 'arch=compute_100a,code=sm_100a',
              '-lineinfo', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found')


def _sources_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cu', '.cuh', '.h')):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(os.path.abspath(__file__)))


def _compile(args):
    src, obj, defs, log = args
    cmd = [_nvcc()] + NVCC_FLAGS + defs + (['-DBDOF_ALT=%d' % ALT] if ALT in (1, 2) else []) + (['-DBDOF_PHASE_TIMING'] if ALT == 9 else []) + (['-DBDOF_NO_STAGGER'] if ALT in (3, 5) else []) + (['-DBDOF_NO_ROWPF'] if ALT in (4, 5) else []) + (['-DBDOF_NO_ROW_TILE_PREFETCH'] if ALT in (6, 8) else []) + (['-DBDOF_NO_COL_TILE_PREFETCH'] if ALT in (7, 8) else []) + (['-DBDOF_TGROUP=4'] if ALT == 10 else []) + ['-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, 'w') as f:
        f.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s' % (obj, r.stderr[-4000:]))
    return obj


def build_lib(force=False, verbose=True):
    """Compile every CUDA source for sm_100a and link libbdof.so.  Returns the library path."""
    stamp = _sources_mtime()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= stamp:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    jobs = [(os.path.join(CSRC, 'bdof.cu'), os.path.join(OBJ, 'bdof.o'), [], os.path.join(OBJ, 'bdof.log'))]
    jobs.append((os.path.join(CSRC, 'dpexchange.cu'), os.path.join(OBJ, 'dpexchange.o'), [], os.path.join(OBJ, 'dpexchange.log')))
    jobs.append((os.path.join(CSRC, 'genericfft.cu'), os.path.join(OBJ, 'genericfft.o'), [], os.path.join(OBJ, 'genericfft.log')))
    jobs.append((os.path.join(CSRC, 'tilehalo.cu'), os.path.join(OBJ, 'tilehalo.o'), [], os.path.join(OBJ, 'tilehalo.log')))
    jobs.append((os.path.join(CSRC, 'resident_inst.cu'), os.path.join(OBJ, 'resident_inst.o'), [], os.path.join(OBJ, 'resident_inst.log')))
    jobs.append((os.path.join(CSRC, 'cluster_inst.cu'), os.path.join(OBJ, 'cluster_inst.o'), [], os.path.join(OBJ, 'cluster_inst.log')))
    for n in SIZES:
        jobs.append((os.path.join(CSRC, 'line_inst.cu'), os.path.join(OBJ, 'line_%d.o' % n), ['-DBDOF_N=%d' % n],
                     os.path.join(OBJ, 'line_%d.log' % n)))
    todo = [j for j in jobs if force or not os.path.exists(j[1]) or os.path.getmtime(j[1]) < stamp]
    if verbose:
        print('[bdof build] compiling %d unit(s) for sm_100a ...' % len(todo), flush=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(_compile, todo))
    cmd = [_nvcc(), '-shared', '-o', LIB] + [j[1] for j in jobs] + ['-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr)
    if verbose:
        print('[bdof build] wrote', LIB, flush=True)
    return LIB


if __name__ == '__main__':
    build_lib(force='--force' in sys.argv)
