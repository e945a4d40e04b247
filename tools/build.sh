#!/bin/bash
# rebuild lib/libci_b200.so (only the stale translation units); extra args: verbose
cd "$(dirname "$0")/.." && python - "$@" <<'PY'
import sys, time
sys.path.insert(0, 'tfp-causalimpact_b200')
from causalimpact_b200 import _build
t = time.time()
_build.build(verbose='verbose' in sys.argv)
print('build ok in %.1fs' % (time.time() - t))
PY
