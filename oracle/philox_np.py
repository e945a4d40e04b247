"""Philox4x32-10 + Box-Muller in NumPy, bit-identical (integers) to
tfp-causalimpact_b200/csrc/ci_common.cuh.  TEST INFRASTRUCTURE ONLY.

Counter-based RNG (Salmon et al., SC'11).  The reference uses TFP's stateless
seeds (causalimpact_lib.py:535-543, 364); an identical stream is impossible
without TFP, so the engine defines its own keyed streams and the oracle
restates them so that GPU and CPU draw the SAME normals.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

RNG_MOMENTUM, RNG_ACCEPT, RNG_LEAPFROG, RNG_SMOOTH, RNG_PREDICT = 1, 2, 3, 4, 5


def philox4x32(seed, c0, c1, c2, c3):
  """All counters broadcastable integer arrays; returns 4 uint32 arrays."""
  c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & MASK
                                         for c in (c0, c1, c2, c3)))
  k0 = int(seed) & 0xFFFFFFFF
  k1 = (int(seed) >> 32) & 0xFFFFFFFF
  for _ in range(10):
    p0 = M0 * c0
    p1 = M1 * c2
    n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
    n1 = p1 & MASK
    n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
    n3 = p0 & MASK
    c0, c1, c2, c3 = n0, n1, n2, n3
    k0 = (k0 + W0) & 0xFFFFFFFF
    k1 = (k1 + W1) & 0xFFFFFFFF
  return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(x):
  return ((x >> np.uint32(8)).astype(np.float64) + 0.5) * 2.0 ** -24


def box_muller(x0, x1):
  u, w = u01(x0), u01(x1)
  rad = np.sqrt(-2.0 * np.log(u))
  return rad * np.cos(2.0 * np.pi * w), rad * np.sin(2.0 * np.pi * w)


def _c1(stream, ident):
  return np.uint64(stream) | ((np.asarray(ident, dtype=np.uint64) >> np.uint64(32)) << np.uint64(8))


def predict_normals(seed, draw_id, T):
  """(z_smooth[T], z_pred[T]) of draw `draw_id`: one Philox call per 2 steps."""
  npair = (T + 1) // 2
  x0, x1, x2, x3 = philox4x32(seed, np.uint64(draw_id) & MASK, _c1(RNG_SMOOTH, draw_id),
                              np.arange(npair), 0)
  a0, a1 = box_muller(x0, x1)
  b0, b1 = box_muller(x2, x3)
  zs = np.empty(2 * npair); zp = np.empty(2 * npair)
  zs[0::2], zp[0::2] = a0, a1
  zs[1::2], zp[1::2] = b0, b1
  return zs[:T], zp[:T]


def momentum_normals(seed, chain_id, it, dim):
  ng = (dim + 3) // 4
  x0, x1, x2, x3 = philox4x32(seed, np.uint64(chain_id) & MASK, _c1(RNG_MOMENTUM, chain_id), it,
                              np.arange(ng))
  a0, a1 = box_muller(x0, x1)
  b0, b1 = box_muller(x2, x3)
  z = np.stack([a0, a1, b0, b1], axis=1).reshape(-1)
  return z[:dim]


def accept_uniform(seed, chain_id, it):
  x0, _, _, _ = philox4x32(seed, np.uint64(chain_id) & MASK, _c1(RNG_ACCEPT, chain_id), it, 0)
  return float(u01(x0))


def leapfrog_count(seed, it, max_leapfrog):
  x0, _, _, _ = philox4x32(seed, 0, RNG_LEAPFROG, it, 0)
  return 1 + int((int(x0) * int(max_leapfrog)) >> 32)
