"""In-tree nvcc build of libci_b200.so for sm_100a (no JIT cache, no torch)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)                      # tfp-causalimpact_b200/
CSRC = os.path.join(ROOT, "csrc")
LIB_DIR = os.path.join(ROOT, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libci_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "-shared", "-Xcompiler", "-fPIC", "-cudart", "shared",
    "-Xlinker", "-rpath,/usr/local/cuda/lib64",
]


def _sources():
  return sorted(
      os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))
  ) + [os.path.join(os.path.dirname(ROOT), "include", "ci_b200.h")]


def is_stale() -> bool:
  if not os.path.exists(LIB_PATH):
    return True
  t = os.path.getmtime(LIB_PATH)
  return any(os.path.getmtime(s) > t for s in _sources() if os.path.exists(s))


def build(force: bool = False, verbose: bool = False) -> str:
  """Compile csrc/*.cu into lib/libci_b200.so.  Needs nvcc, not a GPU."""
  if not force and not is_stale():
    return LIB_PATH
  nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
  if not os.path.exists(nvcc):
    raise RuntimeError("nvcc not found: cannot build libci_b200.so")
  os.makedirs(LIB_DIR, exist_ok=True)
  cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
      "-o", LIB_PATH, os.path.join(CSRC, "ci_abi.cu")]
  res = subprocess.run(cmd, capture_output=True, text=True)
  if res.returncode != 0:
    raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
  if verbose:
    print(res.stderr)
  return LIB_PATH
