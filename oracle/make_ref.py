"""Installs the UNMODIFIED reference (TEDEouS, /root/reference) into oracle/_ref/ - test / baseline infrastructure.

    python oracle/make_ref.py          # build container only: /root/reference does not exist on the GPU box

oracle/_ref/ is git-ignored (no reference sources enter the history) but not gpurun-ignored, so the installed
package travels to the GPU box, where `bench.py --impl reference` and the `cpu_baseline` leg time the reference's
own `Solution.evaluate(); loss.backward()` on the host cores (`cpu_baseline.kind = "reference"`).  The install is
`pip install --no-index --no-deps --target oracle/_ref` of a /tmp copy of the tree (the source tree is read-only and
setuptools writes build/ and *.egg-info into it).  The reference's optional imports that this image lacks (SALib,
matplotlib, seaborn - plotting / sensitivity callbacks, not on the hot path) are stubbed by `import_reference()`.
"""
import os
import shutil
import subprocess
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference'
REF_DST = os.path.join(HERE, '_ref')


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DST, 'tedeous'))


def install(force: bool = False) -> bool:
    """-> True when oracle/_ref/tedeous exists afterwards."""
    if available() and not force:
        return True
    if not os.path.isdir(REF_SRC):
        return False
    tmp = tempfile.mkdtemp(prefix='tedeous_ref_')
    try:
        src = os.path.join(tmp, 'src')
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns('.git', 'docs', 'examples', 'tutorials',
                                                                    'landscape_visualization'))
        os.makedirs(REF_DST, exist_ok=True)
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--quiet',
               '--upgrade', '--target', REF_DST, src]
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=src)
        if res.returncode != 0 or not available():
            sys.stderr.write(res.stdout + res.stderr)
            # same result as the wheel's purelib payload: the package directory itself
            shutil.copytree(os.path.join(REF_SRC, 'tedeous'), os.path.join(REF_DST, 'tedeous'), dirs_exist_ok=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return available()


def import_reference():
    """Puts oracle/_ref first on sys.path (stubs for the absent plotting / sensitivity packages) and returns the
    `tedeous` package; raises ImportError when the reference was not installed."""
    if not available():
        raise ImportError('oracle/_ref/tedeous is absent: run `python oracle/make_ref.py` in the build container')
    for name in ['SALib', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.cm', 'seaborn']:
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules['SALib'], 'ProblemSpec'):
        sys.modules['SALib'].ProblemSpec = object
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import tedeous  # noqa: F401
    return tedeous


if __name__ == '__main__':
    ok = install(force='--force' in sys.argv)
    print('oracle/_ref:', 'installed' if ok else 'unavailable (no /root/reference here)')
