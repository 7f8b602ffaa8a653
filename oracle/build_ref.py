"""Compile the REFERENCE's own two CUDA ops, from the sources where they lie under /root/reference, into
oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).  TEST INFRASTRUCTURE ONLY.

  /root/reference/stylegan_code_finder/networks/stylegan2/op/fused_bias_act.cpp + fused_bias_act_kernel.cu -> _ref/ref_fused.so
  /root/reference/stylegan_code_finder/networks/stylegan2/op/upfirdn2d.cpp      + upfirdn2d_kernel.cu      -> _ref/ref_upfirdn2d.so

The reference builds these with torch's JIT `load()` at import time; this recipe runs nvcc / g++ on the four files
directly (no reference build system, no source copied into the repo).  The resulting pybind modules expose exactly
`fused_bias_act(...)` / `upfirdn2d(...)` and are used by tests/test_ref_kernels_gpu.py as the GPU oracle for the two
native ops (bit-exact comparison on the B200).  The generator itself (model.py) is Python and cannot travel.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
SRC = '/root/reference/stylegan_code_finder/networks/stylegan2/op'
TARGETS = {'ref_fused': ('fused_bias_act.cpp', 'fused_bias_act_kernel.cu'),
           'ref_upfirdn2d': ('upfirdn2d.cpp', 'upfirdn2d_kernel.cu')}


def main():
    if not os.path.isdir(SRC):
        print('reference sources not present; nothing built')
        return 0
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    inc = [f'-I{p}' for p in ce.include_paths('cuda')] + [f'-I{sysconfig.get_paths()["include"]}']
    torch_lib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    abi = f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}'
    for name, (cpp, cu) in TARGETS.items():
        so = os.path.join(OUT, f'{name}.so')
        srcs = [os.path.join(SRC, cpp), os.path.join(SRC, cu)]
        if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(s) for s in srcs):
            print(f'{so} up to date')
            continue
        common = [f'-DTORCH_EXTENSION_NAME={name}', '-DTORCH_API_INCLUDE_EXTENSION_H', abi, '-std=c++17', '-O3'] + inc
        o_cpp, o_cu = os.path.join(OUT, f'{name}_cpp.o'), os.path.join(OUT, f'{name}_cu.o')
        subprocess.run(['g++', '-c', srcs[0], '-o', o_cpp, '-fPIC'] + common, check=True)
        subprocess.run(['/usr/local/cuda/bin/nvcc', '-c', srcs[1], '-o', o_cu, '-gencode', 'arch=compute_100,code=sm_100',
                        '--compiler-options', '-fPIC', '--expt-relaxed-constexpr'] + common, check=True)
        subprocess.run(['g++', '-shared', '-o', so, o_cpp, o_cu, f'-L{torch_lib}', '-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch_cuda',
                        '-ltorch', '-ltorch_python', '-L/usr/local/cuda/lib64', '-lcudart', f'-Wl,-rpath,{torch_lib}'], check=True)
        os.remove(o_cpp); os.remove(o_cu)
        print(f'built {so}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
