"""CPU oracle for the JABD box-geometry hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``jabd_b200``) never does: it has no CPU path at all.

* ``oracle.oracle``      numpy-facing wrapper of ``jabd_oracle.c`` (scalar C
                         restatement; the checker of record on the GPU box).
* ``oracle.torch_port``  op-for-op torch-CPU port of the reference's eager code
                         path (same temporaries, same third-party
                         ``torchvision.ops.nms``); used as the timed CPU baseline.

Parity is pinned against outputs of the imported reference, see
``tests/golden/make_golden.py`` and ``tests/test_oracle_golden.py``.
"""
