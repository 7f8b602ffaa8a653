"""In-tree build of lib/libsis_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build_library(force: bool = False, verbose: bool = False) -> str:
    csrc = os.path.join(_HERE, 'csrc')
    os.makedirs(os.path.join(_HERE, 'lib'), exist_ok=True)
    cmd = ['make', '-C', csrc, '-j', str(os.cpu_count() or 4)]
    if force:
        subprocess.run(['make', '-C', csrc, 'clean'], check=True, capture_output=not verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('building libsis_b200.so failed:\n' + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout)
    return os.path.join(_HERE, 'lib', 'libsis_b200.so')
