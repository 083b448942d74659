"""CPU oracle for the birda front-end / post-inference hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product path (``birda_b200``) never does and
fails loudly when its CUDA library is missing.

It is a function-by-function restatement (numpy, f32-faithful) of the reference's
algorithm for SURVEY.md §8a rows A1-A10; every function cites the reference
file:line it follows (paths relative to ``/root/reference``).

Parity pinning status:

* segmentation / sizing / time stamps / downmix / mask / sort / date math follow
  in-tree Rust source and are pinned by the reference's own unit-test tables,
  ported in ``tests/test_oracle_*.py``;
* the resampler (third-party ``rubato`` 4.0.0, not vendored in the reference
  tree) and activation/top-k (third-party ``birdnet-onnx`` 2.0.0-rc.16) are
  restated from their published algorithms and validated only against the
  reference's *property* tests (``src/audio/resample.rs:240-385``):
  **sample-level parity of those two steps is unpinned**.
"""

from . import frontend, post, rules  # noqa: F401
