"""In-tree build of the CUDA extension (libble_b200.so) for sm_100a.

    python -m balloon_learning_environment_b200._build

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the tree.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_PATH = os.environ.get('BLE_B200_LIB') or os.path.join(PKG_DIR, 'libble_b200.so')
SOURCES = [os.path.join(CSRC, 'ble_engine.cu')]
HEADERS = [os.path.join(CSRC, f) for f in ('ble_physics.cuh', 'ble_wind.cuh', 'ble_features.cuh',
                                          'ble_feature_kernels.cuh', 'ble_decoder.cuh')] + [
           os.path.join(PKG_DIR, '..', 'include', 'ble_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--shared', '-Xcompiler', '-fPIC']
LINK_FLAGS = ['-lcublasLt', '-Xlinker', '-rpath=/usr/local/cuda/lib64']


def find_nvcc():
  for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError('nvcc not found: cannot build libble_b200.so (set NVCC=/path/to/nvcc)')


def is_stale():
  if not os.path.exists(LIB_PATH):
    return True
  built = os.path.getmtime(LIB_PATH)
  return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
  """Compiles csrc/*.cu -> libble_b200.so if missing or out of date.  Returns the library path."""
  if not force and not is_stale():
    return LIB_PATH
  cmd = [find_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH] + SOURCES + LINK_FLAGS
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
  if verbose:
    sys.stderr.write(proc.stderr)
  return LIB_PATH


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
