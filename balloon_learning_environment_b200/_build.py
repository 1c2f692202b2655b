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
SOURCES = [os.path.join(CSRC, f) for f in ('ble_engine.cu', 'ble_step_fused.cu', 'ble_learner.cu', 'ble_dense.cu')]
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))) + [
           os.path.join(PKG_DIR, '..', 'include', 'ble_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--shared', '-Xcompiler', '-fPIC']
LINK_FLAGS = ['-lcublasLt', '-Xlinker', '-rpath=/usr/local/cuda/lib64']


def find_nvcc():
  for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError('nvcc not found: cannot build libble_b200.so (set NVCC=/path/to/nvcc)')


OBJ_DIR = os.path.join(PKG_DIR, 'build')


def _newer_than(target, deps):
  if not os.path.exists(target):
    return True
  built = os.path.getmtime(target)
  return any(os.path.getmtime(p) > built for p in deps)


def is_stale():
  return _newer_than(LIB_PATH, SOURCES + HEADERS)


def build(force=False, verbose=False):
  """Compiles csrc/*.cu -> build/*.o -> libble_b200.so, redoing only what is out of date.  Returns the library path."""
  if not force and not is_stale():
    return LIB_PATH
  nvcc = find_nvcc()
  os.makedirs(OBJ_DIR, exist_ok=True)
  compile_flags = [f for f in NVCC_FLAGS if f != '--shared']
  objects, jobs = [], []
  for src in SOURCES:
    obj = os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + '.o')
    objects.append(obj)
    if not force and not _newer_than(obj, [src] + HEADERS):
      continue
    cmd = [nvcc] + compile_flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, src]
    jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))   # one nvcc per source, concurrently
  for cmd, proc in jobs:
    out, errtxt = proc.communicate()
    if proc.returncode != 0:
      raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + out + errtxt)
    if verbose:
      sys.stderr.write(errtxt)
  cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '--shared', '-o', LIB_PATH] + objects + LINK_FLAGS
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    raise RuntimeError('nvcc link failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
  return LIB_PATH


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
