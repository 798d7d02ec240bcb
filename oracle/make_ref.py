"""Recipe for ``oracle/_ref/`` -- the UNMODIFIED reference hot path, made importable next to this repo.

    python oracle/make_ref.py            (build container only: needs /root/reference)

The reference (uber-research/vargp) is pure Python with no build system; its hot path is four files that import
nothing but torch (var_gp/{vargp,kernels,gp_utils,likelihoods}.py).  This script copies exactly those files,
byte for byte, into the git-ignored ``oracle/_ref/refvargp/`` (a different package name: ``var_gp`` at the repo root
is the drop-in shim of the product).  ``oracle/_ref/`` is NOT gpurun-ignored, so the copy travels to the GPU box like
a built ``.so``, where ``bench.py --impl reference`` and the ``cpu_baseline`` leg time it on the host cores
(``cpu_baseline.kind = "reference"``).  No reference source enters the git history.

TEST / BENCH INFRASTRUCTURE ONLY: nothing under ``vargp_b200/`` imports it (oracle/ref_runner.py is the one loader).
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('VARGP_REFERENCE', '/root/reference')
FILES = ('vargp.py', 'kernels.py', 'gp_utils.py', 'likelihoods.py', 'vargp_retrain.py')


def main():
  src = os.path.join(REF, 'var_gp')
  if not os.path.isdir(src):
    print(f'make_ref: {src} not present (GPU box?) -- keeping whatever oracle/_ref already holds')
    return 0
  dst = os.path.join(HERE, '_ref', 'refvargp')
  os.makedirs(dst, exist_ok=True)
  lines = []
  for f in FILES:
    shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    lines.append(f'{hashlib.sha256(open(os.path.join(dst, f), "rb").read()).hexdigest()}  var_gp/{f}')
  open(os.path.join(dst, '__init__.py'), 'w').close()
  open(os.path.join(HERE, '_ref', 'SHA256SUMS'), 'w').write('\n'.join(lines) + '\n')
  print(f'make_ref: copied {len(FILES)} files of {src} -> {dst}')
  return 0


if __name__ == '__main__':
  sys.exit(main())
