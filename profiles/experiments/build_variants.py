"""Build A/B variants of libp3d.so that differ only in -D switches of the fused renderers (render_fused_ws.cu: P3D_WS_*,
render_fused_ws3.cu: P3D_W3_* - see the #ifndef blocks at their tops).

    python profiles/experiments/build_variants.py name:-DX=1,-DY=2 [name2:...]   ->  build/variants/libp3d_<name>.so

Select one at run time with P3D_LIBP3D=<path> (panic3d_b200/_lib.py).  build/ is git-ignored but travels to the GPU box."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import panic3d_b200  # noqa: E402,F401
from panic3d_b200 import _build  # noqa: E402


def main():
    _build.build()                                        # make sure every object is fresh
    objdir = os.path.join(ROOT, 'gpurun_out', '.obj')
    out = os.path.join(ROOT, 'build', 'variants')
    os.makedirs(out, exist_ok=True)
    fused = ('render_fused_ws.cu', 'render_fused_ws3.cu')
    others = [os.path.join(objdir, os.path.basename(s)[:-3] + '.o') for s in _build.sources() if not s.endswith(fused)]
    procs = []
    for spec in sys.argv[1:]:
        name, _, defs = spec.partition(':')
        objs = []
        for f in fused:
            obj = os.path.join(out, f'{f[:-3]}_{name}.o')
            cmd = ['nvcc'] + [x for x in _build.NVCC_FLAGS if x != '-shared'] + [d for d in defs.split(',') if d] + ['-c', os.path.join(_build.CSRC, f), '-o', obj]
            procs.append((name, subprocess.Popen(cmd, stderr=subprocess.DEVNULL)))
            objs.append(obj)
        procs.append((name, objs))
    pending = {}
    for item in procs:
        if isinstance(item[1], list):
            name, objs = item
            lib = os.path.join(out, f'libp3d_{name}.so')
            subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', lib] + others + objs)
            for o in objs:
                os.remove(o)
            print(lib)
        else:
            assert item[1].wait() == 0, item[0]


if __name__ == '__main__':
    main()
