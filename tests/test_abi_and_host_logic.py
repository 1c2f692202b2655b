"""CPU-side checks: the C-ABI library builds, loads and exports every symbol of include/ble_b200.h;
the host logic rejects bad input; the product never reaches for the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
  from balloon_learning_environment_b200 import _lib
  lib = _lib.load()
  header = open(os.path.join(ROOT, 'include', 'ble_b200.h')).read()
  declared = sorted(set(re.findall(r'\b(ble_[a-z0-9_]+)\s*\(', header)))
  assert len(declared) >= 15
  for name in declared:
    assert hasattr(lib, name), f'{name} declared in include/ble_b200.h but not exported'
  assert set(declared) == set(_lib.EXPORTS)


def test_header_is_plain_c(tmp_path):
  """The boundary is a C ABI: include/ble_b200.h must compile as C99 (no C++, CUDA or torch types)."""
  import shutil
  import subprocess
  gcc = shutil.which('gcc')
  if gcc is None:
    pytest.skip('gcc not found')
  src = tmp_path / 'abi.c'
  src.write_text('#include "ble_b200.h"\nint main(void) { ble_config c; ble_state_soa s; (void)c; (void)s; return 0; }\n')
  proc = subprocess.run([gcc, '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-fsyntax-only',
                         '-I', os.path.join(ROOT, 'include'), str(src)], capture_output=True, text=True)
  assert proc.returncode == 0, proc.stderr


def test_c_program_links_and_calls_the_library(tmp_path):
  """A plain C caller (what a cgo / JNI / FFI binding boils down to): links libble_b200.so, rejects bad
  arguments, and - without a GPU - gets BLE_ERR_CUDA with the 'no CPU fallback' message."""
  import shutil
  import subprocess
  from balloon_learning_environment_b200 import _build
  gcc = shutil.which('gcc')
  if gcc is None:
    pytest.skip('gcc not found')
  lib_dir = os.path.dirname(_build.build())
  src = tmp_path / 'caller.c'
  src.write_text(r"""
#include <stdio.h>
#include <string.h>
#include "ble_b200.h"
int main(void) {
  ble_config cfg; memset(&cfg, 0, sizeof cfg);
  cfg.precision = BLE_PRECISION_FP32; cfg.wind_model = BLE_WIND_GRID; cfg.enable_noise = 1; cfg.field_layout = BLE_LAYOUT_X64;
  ble_handle* h = NULL;
  if (ble_create(0, 0, &cfg, &h) != BLE_ERR_INVALID_ARGUMENT) return 10;
  if (ble_create(0, 64, NULL, &h) != BLE_ERR_INVALID_ARGUMENT) return 11;
  if (ble_step(NULL, NULL, NULL, NULL, NULL, NULL) != BLE_ERR_INVALID_ARGUMENT) return 12;
  if (ble_num_envs(NULL) != 0 || ble_destroy(NULL) != BLE_ERR_INVALID_ARGUMENT) return 13;
  int rc = ble_create(0, 64, &cfg, &h);
  printf("%d|%s\n", rc, ble_last_error(NULL));
  if (rc == BLE_OK) { long long n = (long long)ble_num_envs(h); ble_destroy(h); return n == 64 ? 0 : 14; }
  return rc == BLE_ERR_CUDA ? 0 : 15;
}
""")
  exe = tmp_path / 'caller'
  proc = subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe),
                         '-L', lib_dir, '-lble_b200', f'-Wl,-rpath,{lib_dir}'], capture_output=True, text=True)
  assert proc.returncode == 0, proc.stderr
  run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
  assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
  rc, msg = run.stdout.strip().split('|', 1)
  assert int(rc) == 0 or 'no CPU fallback' in msg


def test_create_without_gpu_fails_loudly():
  import torch
  if torch.cuda.is_available():
    pytest.skip('a GPU is present')
  from balloon_learning_environment_b200 import _lib, batched_env
  lib = _lib.load()
  cfg = _lib.BleConfig(0, 0, 1, 0)
  h = ctypes.c_void_p()
  rc = lib.ble_create(0, 16, ctypes.byref(cfg), ctypes.byref(h))
  assert rc == -2 and b'no CPU fallback' in lib.ble_last_error(None)
  assert lib.ble_create(0, 0, ctypes.byref(cfg), ctypes.byref(h)) == -1       # invalid argument
  with pytest.raises(_lib.BleError):
    batched_env.BatchedBalloonArena(4)


def test_stateless_learner_entry_points_reject_bad_arguments():
  """The stateless entry points validate their arguments before they touch the device: null pointers, pitches that are
  not TMA-able, inconsistent modes -> BLE_ERR_INVALID_ARGUMENT (no GPU needed to get there)."""
  import numpy as np
  from balloon_learning_environment_b200 import _lib
  lib = _lib.load()
  buf = np.zeros(4096, np.float32)
  base = buf.ctypes.data + (-buf.ctypes.data) % 16                 # a 16-byte aligned host address: never dereferenced
  p = ctypes.c_void_p(base)
  null = ctypes.c_void_p(0)
  inv = -1                                                         # BLE_ERR_INVALID_ARGUMENT (include/ble_b200.h)

  def dense(a=p, lda=32, b=p, ldb=32, m=8, n=8, k=32, mode=0, aux=p, ld_aux=0, d=p, ldd=8, dt=null, ldt=0, split=1,
            bits=null, ld_bits=0):
    return lib.ble_dense_tf32(a, lda, b, ldb, m, n, k, mode, aux, ld_aux, d, ldd, dt, ldt, split, bits, ld_bits, null)

  assert dense(a=null) == inv and dense(b=null) == inv
  assert dense(lda=30) == inv                                      # pitch not a multiple of 4 floats (16-byte TMA rows)
  assert dense(lda=16) == inv                                      # pitch shorter than K
  assert dense(a=ctypes.c_void_p(base + 4)) == inv                 # base not 16-byte aligned
  assert dense(mode=5) == inv and dense(mode=-1) == inv
  assert dense(aux=null) == inv                                    # bias missing
  assert dense(d=null) == inv                                      # no output at all
  assert dense(ldd=4) == inv                                       # output pitch shorter than N
  assert dense(split=2) == inv                                     # split-K only with the accumulating modes
  assert dense(mode=3, d=p, dt=null) == inv and dense(mode=3, d=p, dt=p, ldt=8) == inv   # accumulates into dt only
  assert dense(mode=2, aux=null, bits=null) == inv                 # a mask from somewhere
  assert dense(mode=1, bits=p, ld_bits=0) == inv                   # mask pitch shorter than ceil(n / 32) words
  assert dense(mode=4, d=null, dt=p, ldt=8, lda=4) == inv          # MN-major: pitch must cover M
  assert lib.ble_transpose_f32(null, 8, 8, 8, p, 8, null) == inv and lib.ble_transpose_f32(p, 4, 8, 8, p, 8, null) == inv
  assert lib.ble_row_sum_f32(p, 8, 8, 8, null, 0, null) == inv and lib.ble_row_sum_f32(p, 4, 8, 8, p, 0, null) == inv


def test_row_order_matches_header_enums():
  from balloon_learning_environment_b200 import _lib
  header = open(os.path.join(ROOT, 'include', 'ble_b200.h')).read()
  f_names = re.findall(r'BLE_F_([A-Z_]+)', header.split('BLE_NUM_F')[0].split('enum {')[-1])
  i_names = re.findall(r'BLE_I_([A-Z_]+)', header.split('BLE_NUM_I')[0].split('enum {')[-1])
  assert [n.lower() for n in f_names] == list(_lib.F_ROWS)
  assert [n.lower() for n in i_names] == list(_lib.I_ROWS)


def test_product_package_does_not_import_the_oracle():
  pkg = os.path.join(ROOT, 'balloon_learning_environment_b200')
  for dirpath, _, files in os.walk(pkg):
    for fn in files:
      if fn.endswith(('.py', '.cu', '.cuh', '.h')):
        text = open(os.path.join(dirpath, fn)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle', text, re.M), fn
        assert not re.search(r'#include\s+".*hostemu', text), fn


def test_eval_suites_and_result_schema():
  """eval/suites.py:39-63, eval/eval.py:121-124 (sharding) and eval/eval_lib.py:33-56 (JSON schema)."""
  import json
  import os
  from balloon_learning_environment_b200 import eval_lib, suites
  assert suites.available_suites() == ['big_eval', 'medium_eval', 'small_eval', 'tiny_eval', 'micro_eval']
  big = suites.get_eval_suite('big_eval')
  assert len(big.seeds) == 10_000 and big.max_episode_length == 960 and big.seeds[:3] == [0, 1, 2]
  assert suites.get_eval_suite('micro_eval').seeds == [0]
  parts = [suites.shard(big, i, 7) for i in range(7)]
  assert sum((p.seeds for p in parts), []) == big.seeds
  with pytest.raises(ValueError):
    suites.get_eval_suite('nope')
  r = eval_lib.EvaluationResult(seed=3, cumulative_reward=1.5, time_within_radius=0.25, out_of_power=False,
                                envelope_burst=False, zeropressure=True, final_timestep=2,
                                flight_path=[eval_lib.SimpleBalloonState(1.0, 2.0, 9000.0, 500.0, 180.0, 0.9)] * 2)
  mine = json.loads(eval_lib.results_to_json([r]))
  here = os.path.dirname(os.path.abspath(__file__))
  with open(os.path.join(here, 'golden', 'eval.json')) as f:
    ref = json.load(f)
  assert list(mine[0].keys()) == list(ref[0].keys())
  assert list(mine[0]['flight_path'][0].keys()) == list(ref[0]['flight_path'][0].keys())
  assert {k: type(v) for k, v in mine[0].items()} == {k: type(v) for k, v in ref[0].items()}
  assert 'seed=3' in str(r)
