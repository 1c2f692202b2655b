"""Decoded wind fields of the reference's REAL checkpoint (this container only) -> tests/golden/decoder_real.npz.

    python -m tests.golden.tier0.make_decoder_golden

models/offlineskies22_decoder.msgpack (25.9 MB) is run through the reference's own flax module
(generative/vae.py:134-186, `vae.Decoder().apply(params, z)`) under the Tier-0 stubs for three seeded latents; the
latents and the three [21,21,10,9,2] fields are the golden.  The checkpoint itself is not committed: the build step
(`__graft_entry__.build()`) places a copy under oracle/_ref/ (git-ignored, travels to the GPU box), where
tests/test_gpu_parity.py::test_decoder_on_the_real_checkpoint loads it with the product's own msgpack reader.
"""
import os

import tests.golden.tier0.boot as boot  # noqa: F401

import numpy as np
from balloon_learning_environment.generative import vae
from flax import serialization

OUT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKPOINT = os.path.join(boot.REFERENCE_ROOT, 'balloon_learning_environment', 'models', 'offlineskies22_decoder.msgpack')


def main():
  real = serialization.msgpack_restore(open(CHECKPOINT, 'rb').read())
  z = np.random.default_rng(2022).standard_normal((3, 64)).astype(np.float32)
  fields = np.stack([np.asarray(vae.Decoder().apply(real, zi), np.float32) for zi in z])
  assert fields.shape == (3, 21, 21, 10, 9, 2)
  np.savez_compressed(os.path.join(OUT, 'decoder_real.npz'), latents=z, fields=fields,
                      checkpoint_bytes=np.int64(os.path.getsize(CHECKPOINT)))
  print('wrote', fields.shape, 'max |wind|', float(np.abs(fields).max()), 'mean |wind|', float(np.abs(fields).mean()))


if __name__ == '__main__':
  main()
