"""TEST INFRASTRUCTURE ONLY -- CPU restatement of fVDB's sparse-convolution hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The product path
(``fvdb-core_b200/``) never imports this package and fails loudly when its CUDA library is absent.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here against
(a) the known-answer tests the reference holds for this path (``tests/unit/test_conv_semantics.py``,
``src/tests/GatherScatterDefaultConvTest.cu:191-255``) and (b) golden vectors produced by running the
reference's own independent oracle (``fvdb/utils/tests/convolution_semantics_oracle.py``) in the build
container -- see ``tests/golden/make_golden.py``.  The reference's compiled implementation
(``_fvdb_cpp``) cannot be built here (NanoVDB / CUTLASS are fetched at configure time; no network), so
``oracle/_ref`` does not exist; DESIGN.md records that.
"""

from .conv_oracle import *  # noqa: F401,F403
from . import pool_oracle  # noqa: F401,E402  (pooling / refinement restatement; pinned against dense torch pooling in tests/test_oracle_golden.py)
