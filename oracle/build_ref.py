"""Build the UNMODIFIED-SEMANTICS reference CUDA extension into oracle/_ref/ (test infrastructure only).

This is checker infrastructure, not product code: only tests/, __graft_entry__.smoke() and
bench.py's reference/cpu_baseline legs may load what this script produces.

What it does
------------
The reference's ray-march ops live in /root/reference/lib/cuda/render_utils.cpp and
render_utils_kernel.cu (13 pybind functions, render_utils.cpp:170-184).  They do not compile
against torch 2.11 as shipped: the 13 `AT_DISPATCH_FLOATING_TYPES(x.type(), ...)` sites in
render_utils_kernel.cu need a `ScalarType`, and `.type()` returns `DeprecatedTypeProperties`
(SURVEY.md section 0).  This recipe

  1. copies the two source files from where they lie under /root/reference into a scratch
     directory under /tmp (never into the repository),
  2. rewrites `.type()` -> `.scalar_type()` inside the AT_DISPATCH calls (a mechanical,
     semantics-preserving token patch; nothing else is touched),
  3. compiles them for sm_100a with torch.utils.cpp_extension (no fast-math, nvcc defaults,
     i.e. the flags the reference's own JIT `load(...)` call would use, lib/dvgo.py:14-19),
  4. leaves ONLY the built shared object in oracle/_ref/render_utils_cuda.so.

oracle/_ref/ is git-ignored (binary, derived from reference sources) but travels to the GPU box
with the gpurun snapshot, where tests use it as the GPU-side reference for the op-level
restatement in oracle/render_utils_ref.c.  /root/reference does not exist on the GPU box, so this
script is a no-op there (it keeps whatever prebuilt .so travelled with the snapshot).
"""
import os
import re
import shutil
import sys
import tempfile

REF = '/root/reference/lib/cuda'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_ref')
OUT_SO = os.path.join(OUT_DIR, 'render_utils_cuda.so')
OUT_UB360 = os.path.join(OUT_DIR, 'ub360_utils_cuda.so')      # cumdist_thres of DirectContractedVoxGO
MODULES = {
    'render_utils_cuda': (('render_utils.cpp', 'render_utils_kernel.cu'), OUT_SO),
    'ub360_utils_cuda': (('ub360_utils.cpp', 'ub360_utils_kernel.cu'), OUT_UB360),
    # grid-maintenance steps either side of the path during training
    'total_variation_cuda': (('total_variation.cpp', 'total_variation_kernel.cu'), os.path.join(OUT_DIR, 'total_variation_cuda.so')),
    'adam_upd_cuda': (('adam_upd.cpp', 'adam_upd_kernel.cu'), os.path.join(OUT_DIR, 'adam_upd_cuda.so')),
}


def _build_one(name, files, out_so, verbose):
    work = tempfile.mkdtemp(prefix='k4_refbuild_')
    srcs = []
    for fn in files:
        with open(os.path.join(REF, fn)) as f:
            text = f.read()
        text = re.sub(r'(AT_DISPATCH_FLOATING_TYPES\(\s*[A-Za-z_0-9]+)\.type\(\)', r'\1.scalar_type()', text)
        dst = os.path.join(work, fn)
        with open(dst, 'w') as f:
            f.write(text)
        srcs.append(dst)
    from torch.utils.cpp_extension import load
    bdir = os.path.join(work, 'build')
    os.makedirs(bdir)
    load(name=name, sources=srcs, build_directory=bdir, verbose=verbose, is_python_module=False)
    shutil.copyfile(os.path.join(bdir, name + '.so'), out_so)
    shutil.rmtree(work, ignore_errors=True)


PRETRAINED_SRC = '/root/reference/pretrained/RealESRNet_x4plus.pth'
PRETRAINED = os.path.join(OUT_DIR, 'RealESRNet_x4plus.pth')     # the only real-weights fixture the reference ships


def build(force=False, verbose=False):
    if not os.path.isdir(REF):
        return OUT_SO if os.path.exists(OUT_SO) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    # the shipped VC-Decoder initialisation (run_sr.py:663 loads it strict=False): a data asset, git-ignored
    # like the built extensions, copied so that the GPU-box decoder tests can load real weights
    if os.path.exists(PRETRAINED_SRC) and (force or not os.path.exists(PRETRAINED)):
        shutil.copyfile(PRETRAINED_SRC, PRETRAINED)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    os.environ.setdefault('MAX_JOBS', str(os.cpu_count() or 4))
    for name, (files, out_so) in MODULES.items():
        if force or not os.path.exists(out_so):
            _build_one(name, files, out_so, verbose)
    return OUT_SO


if __name__ == '__main__':
    p = build(force='--force' in sys.argv, verbose=True)
    print('oracle/_ref:', p)
