"""Build csrc/ into libb200ns.so with nvcc for sm_100a (in-tree, so the .so travels with gpurun)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ACT_BF16 = os.environ.get('B200NS_ACT', 'fp16').lower() in ('bf16', 'bfloat16')
OUT = os.path.join(HERE, 'libb200ns_bf16.so' if ACT_BF16 else 'libb200ns.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-shared',
              '-Xcompiler', '-fPIC']


def _newest_src():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
               if f.endswith(('.cu', '.cuh', '.h'))) if os.path.isdir(CSRC) else 0


def build(force: bool = False, verbose: bool = False) -> str:
    hdr = os.path.join(os.path.dirname(HERE), 'include', 'b200_noise_search.h')
    newest = max(_newest_src(), os.path.getmtime(hdr) if os.path.exists(hdr) else 0)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    # build into a private temp file and rename atomically: concurrent builders (torchrun ranks) or a process that is
    # dlopen-ing the library never see a half-written .so
    tmp = f'{OUT}.{os.getpid()}.tmp'
    cmd = [nvcc] + NVCC_FLAGS + (['-DB200NS_ACT_BF16'] if ACT_BF16 else []) + (['-Xptxas', '-v'] if verbose else []) + ['-o', tmp, os.path.join(CSRC, 'capi.cu')]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError('nvcc failed building libb200ns.so')
    os.replace(tmp, OUT)
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
