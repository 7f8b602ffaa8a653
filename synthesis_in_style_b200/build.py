"""In-tree build of lib/libsis_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import json
import os
import re
import socket
import subprocess
import time

_HERE = os.path.dirname(os.path.abspath(__file__))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """`make` in csrc/ (object files are rebuilt when their source or any header changed).  Writes
    lib/build_stamp.json: when / where the last build ran and which objects nvcc actually recompiled, so a bench line
    can say whether the library it loaded was compiled by this checkout's build() or travelled prebuilt."""
    csrc = os.path.join(_HERE, 'csrc')
    lib_dir = os.path.join(_HERE, 'lib')
    os.makedirs(lib_dir, exist_ok=True)
    cmd = ['make', '-C', csrc, '-j', str(os.cpu_count() or 4)]
    if force:
        subprocess.run(['make', '-C', csrc, 'clean'], check=True, capture_output=not verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('building libsis_b200.so failed:\n' + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout)
    compiled = sorted(set(re.findall(r'-c (\S+\.cu)', res.stdout)))
    stamp = {'when': time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime()), 'host': socket.gethostname(),
             'forced_clean': bool(force), 'recompiled_sources': compiled, 'relinked': '-shared' in res.stdout,
             'up_to_date': not compiled and '-shared' not in res.stdout}
    try:
        prev_path = os.path.join(lib_dir, 'build_stamp.json')
        if stamp['up_to_date'] and os.path.exists(prev_path):
            with open(prev_path) as f:
                prev = json.load(f)
            stamp['last_compile'] = prev.get('last_compile', {k: prev.get(k) for k in ('when', 'host', 'recompiled_sources')})
        else:
            stamp['last_compile'] = {k: stamp[k] for k in ('when', 'host', 'recompiled_sources')}
        with open(prev_path, 'w') as f:
            json.dump(stamp, f, indent=1)
    except OSError:
        pass
    return os.path.join(lib_dir, 'libsis_b200.so')
