"""Builds libtedeous_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI)."""
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB_DIR = os.path.join(PKG, 'lib')
LIB = os.environ.get('TDB200_LIB') or os.path.join(LIB_DIR, 'libtedeous_b200.so')   # TDB200_LIB: use a prebuilt variant
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '--use_fast_math=false', '-Xptxas', '-v']


def sources():
    return sorted(glob.glob(os.path.join(PKG, 'csrc', '*.cu')))


def needs_build() -> bool:
    if os.environ.get('TDB200_LIB'):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(PKG, 'csrc', '*.cuh')) + glob.glob(os.path.join(ROOT, 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Builds under an exclusive file lock (one process per GPU under torchrun: every rank calls this at import) into a
    private object directory and moves the finished library into place atomically, so no rank can dlopen a half-written
    file.  A prebuilt library is used as it is when nvcc is not available (GPU boxes receive the built .so)."""
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which('nvcc') or ('/usr/local/cuda/bin/nvcc' if os.path.exists('/usr/local/cuda/bin/nvcc') else None)
    if nvcc is None:
        if os.path.exists(LIB):
            return LIB                                 # sources look newer (fresh checkout) but there is no compiler here
        raise RuntimeError('libtedeous_b200.so is not built and nvcc was not found')
    os.makedirs(LIB_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIB_DIR, '.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():        # another rank built it while this one waited
                return LIB
            return _build_locked(nvcc, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(nvcc, verbose):
    def compile_one(src):
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + '.o')
        extra = ['-DTDB_TC_TIMING'] if os.environ.get('TDB200_TC_TIMING_BUILD') else []
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != '--use_fast_math=false'] + extra + \
              ['-I', os.path.join(ROOT, 'include'), '-I', os.path.join(PKG, 'csrc'), '-c', src, '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}')
        with open(obj + '.ptxas.log', 'w') as f:
            f.write(res.stderr)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=8) as ex:      # one nvcc process per translation unit
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB + f'.tmp{os.getpid()}'
    cmd = [nvcc, '-shared', '-o', tmp] + objs + ['-cudart', 'static']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('link failed')
    os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
