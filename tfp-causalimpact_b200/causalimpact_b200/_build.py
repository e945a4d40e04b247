"""In-tree nvcc build of libci_b200.so for sm_100a (no JIT cache, no torch).

csrc/abi_*.cu are independent translation units (each instantiates only the kernels it
launches); they are compiled in parallel into build/*.o and linked into lib/libci_b200.so.
Staleness is decided by CONTENT (sha1 of a unit and of every header it includes,
transitively), not by mtime: a snapshot copied to another box keeps its library.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import json
import os
import re
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)                      # tfp-causalimpact_b200/
CSRC = os.path.join(ROOT, "csrc")
LIB_DIR = os.path.join(ROOT, "lib")
OBJ_DIR = os.path.join(ROOT, "build")
LIB_PATH = os.path.join(LIB_DIR, "libci_b200.so")
STAMP_PATH = LIB_PATH + ".stamp"
HEADER = os.path.join(os.path.dirname(ROOT), "include", "ci_b200.h")

CC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC",
]
LINK_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared",
    "-Xlinker", "-rpath,/usr/local/cuda/lib64", "-ldl",
]
# units whose float32 / float64 kernels go into separate objects (-DCI_ONLY=0 / 1): the slowest
# ones, so that the critical path of a full build is ~25 s instead of ~50 s
SPLIT_BY_DTYPE = ("abi_hmc.cu", "abi_seasonal.cu")

_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def units():
  """[(source, extra flags, object name)] of every translation unit."""
  out = []
  for f in sorted(os.listdir(CSRC)):
    if not (f.startswith("abi_") and f.endswith(".cu")):
      continue
    if f in SPLIT_BY_DTYPE:
      out.append((f, ["-DCI_ONLY=0"], f[:-3] + "_f32.o"))
      out.append((f, ["-DCI_ONLY=1"], f[:-3] + "_f64.o"))
    else:
      out.append((f, [], f[:-3] + ".o"))
  return out


def _deps(path, seen=None):
  """`path` and every file it #includes with quotes, transitively."""
  seen = set() if seen is None else seen
  path = os.path.normpath(path)
  if path in seen or not os.path.exists(path):
    return seen
  seen.add(path)
  for inc in _INC.findall(open(path, encoding="utf-8").read()):
    _deps(os.path.join(os.path.dirname(path), inc), seen)
  return seen


def _digest(src, flags):
  h = hashlib.sha1(" ".join(CC_FLAGS + flags).encode())
  for d in sorted(_deps(os.path.join(CSRC, src))):
    h.update(os.path.relpath(d, CSRC).encode())
    h.update(open(d, "rb").read())
  return h.hexdigest()


def _want():
  return {obj: _digest(src, flags) for src, flags, obj in units()}


def _have():
  try:
    return json.load(open(STAMP_PATH))
  except Exception:   # pylint: disable=broad-except
    return {}


def is_stale() -> bool:
  return not os.path.exists(LIB_PATH) or _have().get("objects") != _want()


def _sources():
  """Every file the library is built from (for tools that list them)."""
  s = set()
  for src, _, _ in units():
    s |= _deps(os.path.join(CSRC, src))
  return sorted(s)


def build(force: bool = False, verbose: bool = False, jobs: int = 0) -> str:
  """Compile csrc/abi_*.cu into lib/libci_b200.so.  Needs nvcc, not a GPU."""
  want = _want()
  if not force and os.path.exists(LIB_PATH) and _have().get("objects") == want:
    return LIB_PATH
  nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
  if not os.path.exists(nvcc):
    raise RuntimeError("nvcc not found: cannot build libci_b200.so")
  os.makedirs(LIB_DIR, exist_ok=True)
  os.makedirs(OBJ_DIR, exist_ok=True)
  obj_stamp_path = os.path.join(OBJ_DIR, "objects.json")
  try:
    obj_have = json.load(open(obj_stamp_path))
  except Exception:   # pylint: disable=broad-except
    obj_have = {}

  def compile_one(unit):
    src, flags, obj = unit
    out = os.path.join(OBJ_DIR, obj)
    if not force and obj_have.get(obj) == want[obj] and os.path.exists(out):
      return obj, ""
    cmd = [nvcc] + CC_FLAGS + flags + (["-Xptxas", "-v"] if verbose else []) + [
        "-c", os.path.join(CSRC, src), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
      raise RuntimeError(f"nvcc failed on {src} {' '.join(flags)}:\n" + res.stdout + res.stderr)
    return obj, res.stderr

  todo = units()
  n = jobs or min(len(todo), os.cpu_count() or 4)
  with cf.ThreadPoolExecutor(max_workers=n) as ex:
    for obj, log in ex.map(compile_one, todo):
      obj_have[obj] = want[obj]
      if verbose and log:
        print(log)
  json.dump(obj_have, open(obj_stamp_path, "w"), indent=1)
  cmd = [nvcc] + LINK_FLAGS + ["-o", LIB_PATH] + [os.path.join(OBJ_DIR, o) for _, _, o in todo]
  res = subprocess.run(cmd, capture_output=True, text=True)
  if res.returncode != 0:
    raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
  json.dump({"objects": want}, open(STAMP_PATH, "w"), indent=1)
  return LIB_PATH
