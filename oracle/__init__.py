"""CPU oracle for the MaxSquare / IW-MaxSquare / Eval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``maxsquareloss_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the CPU arm that is timed *beside* the CUDA path, never as the product.

What it is: a restatement, in NumPy and CPU PyTorch, of the reference's
algorithm for this path (``/root/reference/utils/loss.py:69-119``,
``/root/reference/utils/eval.py:14-124``) plus the call-site prologue
(``graphs/models/deeplab_multi.py:124,128`` bilinear upsample and
``tools/solve_gta5.py:182-183`` softmax).  The reference is pure Python on top
of PyTorch and NumPy, both un-vendored and unpinned third-party dependencies
(README.md:23 "Pytorch(1.0.0)", requirements.txt has no torch; numpy 1.14.6);
this container has torch 2.11.0 and numpy 2.3.5.

Pinning: the reference has no tests or golden vectors of its own.  The oracle
is pinned against outputs of the reference itself, run in the build container
by ``oracle/make_golden.py`` (which imports ``/root/reference`` by file path)
and frozen under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the
restatement against every frozen vector.
"""
